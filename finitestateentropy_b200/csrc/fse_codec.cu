// fse_codec.cu -- batched FSE (tANS) block encode / decode for sm_100a, byte and 16-bit alphabets.
//
// Replaces, per block, the CPU chains
//   FSE_compress2 -> FSE_compress_wksp          lib/fse_compress.c:632-693
//       HIST_count_wksp, FSE_optimalTableLog, FSE_normalizeCount, FSE_writeNCount,
//       FSE_buildCTable_wksp, FSE_compress_usingCTable (2 interleaved states)
//   FSE_decompress -> FSE_decompress_wksp       lib/fse_decompress.c:255-283
//       FSE_readNCount, FSE_buildDTable, FSE_decompress_usingDTable (fast / safe variants)
//   FSE_compressU16 / FSE_decompressU16         lib/fseU16.c:203-251,306-329 (single state)
//
// What the format allows on a GPU: a block is ONE dependent chain on decode
// (both states share one bitstream) and TWO on encode; each step is a table look-up whose index is
// the previous step's result.  Parallelism is therefore across blocks only: one warp per block, the
// chain on lane 0 (decode) / lanes 0-1 (encode), the other lanes cooperating on what is parallel
// (histogram, symbol spreading, table fill, bit packing of the emitted (value,nbBits) pairs).
// Tables live in shared memory (CTable 10 KB, DTable 16 KB at tableLog 12); that footprint bounds
// the number of resident blocks per SM and with it the throughput -- far below the HBM roofline,
// as SURVEY.md section 7.3 anticipates.
#include "common.cuh"
#include "fse_dev.cuh"
#include "bitsrc_dev.cuh"

namespace fseb {
namespace fsek {

constexpr int WARPS = 4;
constexpr int THREADS = 32 * WARPS;
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr u32 WIN = 512;          // bytes of compressed stream staged per warp (decode)

// ---- warp-cooperative symbol spreading ------------------------------------------------------
// Visit v (0..size-1) of the reference walk lands on cell (v*stride) & mask; cells above `high`
// (parked low-probability symbols) are skipped; the r-th accepted visit receives the symbol whose
// cumulative normalized count covers r.  put(cell, symbol) is called exactly once per cell.
// cum: u16[msv+2] exclusive cumulative counts of the positive norms (shared memory).
template <typename Put>
__device__ inline bool warp_spread(const short* norm, const u16* cum, unsigned msv, unsigned tl, Put put)
{
    unsigned const lane = lane_id();
    u32 const size = 1u << tl, mask = size - 1, stride = (size >> 1) + (size >> 3) + 3;
    // parked symbols, in symbol order, from the top cell downwards (tiny: one lane)
    u32 high = size - 1;
    if (lane == 0) for (u32 s = 0; s <= msv; s++) if (norm[s] == -1) put(high--, s);
    high = __shfl_sync(FULL, high, 0);
    u32 const nAccept = cum[msv + 1];
    u32 accepted = 0;
    for (u32 v0 = 0; v0 < size; v0 += 32) {
        u32 const v = v0 + lane;
        u32 const pos = (v * stride) & mask;
        bool const ok = pos <= high;
        u32 const votes = __ballot_sync(FULL, ok);
        u32 const r = accepted + __popc(votes & ((1u << lane) - 1));
        if (ok && r < nAccept) {
            u32 lo = 0, hi = msv + 1;                    // largest s with cum[s] <= r  (skips empty symbols)
            while (hi - lo > 1) { u32 const mid = (lo + hi) >> 1; if (cum[mid] <= r) lo = mid; else hi = mid; }
            put(pos, lo);
        }
        accepted += __popc(votes);
    }
    // the reference requires the walk to close on cell 0 after the last placement
    // (fse_decompress.c:113); with sum(|norm|) == size that is equivalent to accepting exactly nAccept visits.
    return accepted == nAccept;
}

// =================================================================================================
// decode
// =================================================================================================
template <bool WIDE> struct DecCfg;
template <> struct DecCfg<false> { static constexpr unsigned MSV = FSE_MAX_SV, TL = FSE_MAX_TLOG, CELLS = 1u << FSE_MAX_TLOG; };
template <> struct DecCfg<true>  { static constexpr unsigned MSV = U16_MAX_SV, TL = U16_MAX_TLOG, CELLS = 1u << U16_MAX_TLOG; };

template <bool WIDE>
struct alignas(16) DecWarp {
    u32   dt[1 + DecCfg<WIDE>::CELLS];
    short norm[DecCfg<WIDE>::MSV + 1];
    u16   cum[DecCfg<WIDE>::MSV + 3];
    u16   nextOf[DecCfg<WIDE>::MSV + 1];
    alignas(16) u64 win[WIN / 8 + 2];
};

// The compressed stream is consumed through a 512-byte shared-memory window that the whole warp refills
// with one coalesced 16-byte load per lane (the first version let lane 0 fetch its 8-byte container from
// global memory on every reload: with 200+ KB of shared memory in use the L1 is tiny, every reload went
// to L2 -- 128 cycles per symbol).  Lane 0 then reads its container with two aligned LDS.64.
struct WSrc {                       // reader state, meaningful on lane 0
    u64 base;                       // absolute address of stream byte 0
    u64 len; u64 at; u64 w; unsigned used;
    u64 winBase;                    // absolute address (16-aligned) of window byte 0
    const u64* win;                 // shared memory, WIN/8 + 1 words
};
__device__ __forceinline__ u64 ws_ld64(const WSrc& b, u64 at)
{
    u32 const o = (u32)(b.base + at - b.winBase);
    u32 const sh = (o & 7) * 8;
    u64 const lo = b.win[o >> 3];
    if (sh == 0) return lo;
    return (lo >> sh) | (b.win[(o >> 3) + 1] << (64 - sh));
}
// warp-wide: make the window cover [at - margin, at + 8) ; call with `at` broadcast from lane 0
__device__ __forceinline__ void ws_slide(WSrc& b, u64 at, u64* win, bool force)
{
    u64 const lowest = b.base & ~15ull;
    u64 const pos = b.base + at;
    if (!force && (pos >= b.winBase + 64 || b.winBase <= lowest) && pos + 8 <= b.winBase + WIN) return;     // still covered (warp-uniform)
    u64 nb = ((pos + 8 + 15) & ~15ull) - WIN;
    if (nb < lowest || nb > pos) nb = lowest;                       // short streams / wrap
    u64 const endAddr = (b.base + b.len + 15) & ~15ull;             // never read past the 16-byte envelope of the stream
    unsigned const lane = lane_id();
    u64 const a = nb + 16ull * lane;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (a < endAddr) v = __ldg(reinterpret_cast<const uint4*>(a));
    __syncwarp();
    reinterpret_cast<uint4*>(win)[lane] = v;
    if (lane == 0) win[WIN / 8] = 0;
    b.winBase = nb;
    __syncwarp();
}
__device__ inline u64 ws_open(WSrc& b, u64 len)                    // bitstream.h:272-318 ; window already covers the end of the stream
{
    b.len = len; b.at = 0; b.w = 0; b.used = 0;
    if (len < 1) return err(E_SRC_WRONG);
    if (len >= 8) {
        b.at = len - 8; b.w = ws_ld64(b, b.at);
        u32 const last = (u32)(b.w >> 56);
        if (last == 0) return err(E_GENERIC);
        b.used = 8 - hibit(last);
    } else {
        u64 const x = ws_ld64(b, 0) & ((1ULL << (8 * len)) - 1);   // bytes 0..len-1, byte i at bits 8i (as the reference assembles them)
        b.w = x;
        u32 const last = (u32)(x >> (8 * (len - 1))) & 0xFF;
        if (last == 0) return err(E_CORRUPT);
        b.used = 8 - hibit(last);
        b.used += (unsigned)(8 - len) * 8;
    }
    return len;
}
__device__ __forceinline__ u64 ws_read(WSrc& b, unsigned nb)
{
    u64 const mask = nb ? ((1ULL << nb) - 1) : 0;
    u64 const v = (b.w >> ((64u - b.used - nb) & 63u)) & mask;
    b.used += nb; return v;
}
__device__ __forceinline__ u64 ws_read_fast(WSrc& b, unsigned nb)
{
    u64 const v = (b.w << (b.used & 63u)) >> ((64u - nb) & 63u);
    b.used += nb; return v;
}
__device__ __forceinline__ int ws_refill(WSrc& b)                  // bitstream.h:416-440
{
    if (b.used > 64) return SRC_OVER;
    if (b.at >= 8) { b.at -= b.used >> 3; b.used &= 7; b.w = ws_ld64(b, b.at); return SRC_MORE; }
    if (b.at == 0) return b.used < 64 ? SRC_ENDBUF : SRC_DONE;
    u64 nb = b.used >> 3; int st = SRC_MORE;
    if (b.at < nb) { nb = b.at; st = SRC_ENDBUF; }
    b.at -= nb; b.used -= (unsigned)nb * 8; b.w = ws_ld64(b, b.at);
    return st;
}

// ---- fast chain reader -------------------------------------------------------------------------
// The reference's reader is byte-granular: (ptr, bitsConsumed) and a 64-bit container re-read at every reload.
// Its loop control depends only on those two counters, never on the container's value, and bits that lie outside
// the stream can only flow into a state that is never looked up again (after such a read the very next reload
// reports overflow and the loop emits one symbol of the OTHER state and stops).  So the chain below keeps
//   * the two counters, updated exactly as BIT_reloadDStream does (cheap integer ops off the critical path), and
//   * the actual bits in a 64-bit left-aligned register window fed with aligned 32-bit words from the staged
//     shared-memory window (like the Huff0 decoder), so that a symbol costs LDS + 3 dependent ALU ops.
struct Chain {
    u32 hi, lo, vb;                 // window and number of valid bits in it
    u32 noff;                       // byte offset in the staged window of the next (lower) 32-bit word
    u64 at; u32 used;               // the reference's ptr - start and bitsConsumed
};
__device__ __forceinline__ void ch_fill(Chain& c, const u32* win32)
{
    while (c.vb <= 32) {
        u32 const w = c.noff < WIN ? win32[c.noff >> 2] : 0u; c.noff -= 4;   // below the staged envelope only when the stream is exhausted
        u64 W = ((u64)c.hi << 32) | c.lo;
        W |= (u64)w << (32 - c.vb);
        c.hi = (u32)(W >> 32); c.lo = (u32)W; c.vb += 32;
    }
}
__device__ __forceinline__ u32 ch_take(Chain& c, u32 nb)            // nb <= 16, window holds >= nb valid bits
{
    u32 const v = __funnelshift_l(c.hi, 0, nb);                     // top nb bits (0 for nb == 0)
    c.hi = __funnelshift_l(c.lo, c.hi, nb); c.lo = c.lo << nb;
    c.vb -= nb; c.used += nb;
    return v;
}
__device__ __forceinline__ int ch_reload(Chain& c)                  // BIT_reloadDStream on the counters (bitstream.h:416-440)
{
    if (c.used > 64) return SRC_OVER;
    if (c.at >= 8) { c.at -= c.used >> 3; c.used &= 7; return SRC_MORE; }
    if (c.at == 0) return c.used < 64 ? SRC_ENDBUF : SRC_DONE;
    u64 nb = c.used >> 3; int st = SRC_MORE;
    if (c.at < nb) { nb = c.at; st = SRC_ENDBUF; }
    c.at -= nb; c.used -= (u32)nb * 8;
    return st;
}
// warp-wide window maintenance for the chain: keeps >= 64 staged bytes below the next word
__device__ __forceinline__ void ch_slide(WSrc& b, Chain& c, u64* win)
{
    u32 const noff = __shfl_sync(FULL, c.noff, 0);
    u64 const lowest = b.base & ~15ull;
    if (noff >= 64 || b.winBase <= lowest) return;
    u64 const nextA = b.winBase + noff;                             // absolute address of the next word
    u64 nb = ((nextA + 4 + 15) & ~15ull) - WIN;
    if (nb < lowest || nb > nextA) nb = lowest;
    u64 const endAddr = (b.base + b.len + 15) & ~15ull;
    unsigned const lane = lane_id();
    u64 const a = nb + 16ull * lane;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (a < endAddr) v = __ldg(reinterpret_cast<const uint4*>(a));
    __syncwarp();
    reinterpret_cast<uint4*>(win)[lane] = v;
    c.noff = (u32)(nextA - nb);
    b.winBase = nb;
    __syncwarp();
}
// opens the stream on lane 0: end-mark checks of BIT_initDStream (bitstream.h:272-318), counters, first bits
__device__ inline u64 ch_open(Chain& c, const WSrc& b, const u32* win32)
{
    u64 const len = b.len;
    c.hi = c.lo = c.vb = 0; c.at = 0; c.used = 0;
    if (len < 1) return err(E_SRC_WRONG);
    u64 const e = b.base + len;                                     // one past the last byte
    u64 const top4 = (e + 3) & ~3ull;
    c.noff = (u32)(top4 - 4 - b.winBase);
    u32 const lastWord = win32[c.noff >> 2];
    u32 const last = (lastWord >> (8 * (u32)((e - 1) & 3))) & 0xFF;
    if (last == 0) return err(len >= 8 ? E_GENERIC : E_CORRUPT);
    u32 const skip = (u32)(8 * (top4 - e)) + (8 - hibit(last));    // garbage above the stream + zero padding + end mark
    ch_fill(c, win32);
    (void)ch_take(c, skip > 16 ? 16 : skip); if (skip > 16) (void)ch_take(c, skip - 16);
    ch_fill(c, win32);
    if (len >= 8) { c.at = len - 8; c.used = 8 - hibit(last); }
    else { c.at = 0; c.used = 8 - hibit(last) + (u32)(8 - len) * 8; }
    return len;
}

#define FSEB_DSTEP(ST, OUT) do { u32 const cell_ = cells[ST]; u32 const low_ = ch_take(c, cell_ >> 24); \
                                 OUT = (cell_ >> 16) & 0xFF; ST = (cell_ & 0xFFFF) + low_; } while (0)

// FSE_decompress_usingDTable_generic (fse_decompress.c:178-238).  The chain runs on lane 0 in batches of up to 8 loop
// iterations (<= 48 stream bytes); between batches the warp slides the staged window.  Returns on all lanes.
__device__ inline u64 warp_decode_bytes(u8* out, u64 cap, const u8* cSrc, u64 cSize, const u32* dt, u64* win)
{
    unsigned const lane = lane_id();
    unsigned const tl = dt[0] & 0xFFFF;
    const u32* const cells = dt + 1;
    const u32* const win32 = reinterpret_cast<const u32*>(win);
    long long const omax = (long long)cap;
    long long op = 0;
    WSrc b; b.base = reinterpret_cast<u64>(cSrc); b.len = cSize; b.winBase = 0; b.win = win; b.at = 0; b.w = 0; b.used = 0;
    Chain c; c.hi = c.lo = c.vb = 0; c.noff = 0; c.at = 0; c.used = 0;
    u32 s1 = 0, s2 = 0;
    u64 ret = 0; int phase = 0;                                    // 0 = main loop, 1 = tail, 2 = finished
    ws_slide(b, cSize >= 8 ? cSize - 8 : 0, win, true);
    if (lane == 0) {
        u64 const e = ch_open(c, b, win32);
        if (is_err(e)) { ret = e; phase = 2; }
        else {
            s1 = ch_take(c, tl); ch_reload(c); ch_fill(c, win32);
            s2 = ch_take(c, tl); ch_reload(c); ch_fill(c, win32);
        }
    }
    bool const al4 = (reinterpret_cast<u64>(out) & 3) == 0;
    for (;;) {
        phase = __shfl_sync(FULL, phase, 0);
        if (phase == 2) break;
        ch_slide(b, c, win);
        if (lane == 0) {
            if (phase == 0) {
                for (int it = 0; it < 8; it++) {
                    if (!((ch_reload(c) == SRC_MORE) & (op < omax - 3))) { phase = 1; break; }
                    u32 a0, a1, a2, a3;
                    FSEB_DSTEP(s1, a0); FSEB_DSTEP(s2, a1); ch_fill(c, win32);      // <= 24 bits per pair, >= 33 valid after a fill
                    FSEB_DSTEP(s1, a2); FSEB_DSTEP(s2, a3); ch_fill(c, win32);
                    if (al4) *reinterpret_cast<u32*>(out + op) = a0 | (a1 << 8) | (a2 << 16) | (a3 << 24);
                    else { out[op] = (u8)a0; out[op + 1] = (u8)a1; out[op + 2] = (u8)a2; out[op + 3] = (u8)a3; }
                    op += 4;
                }
            } else {                                               // tail (fse_decompress.c:222-235): a handful of symbols
                for (int it = 0; it < 4; it++) {
                    u32 sy;
                    if (op > omax - 2) { ret = err(E_DST_TOO_SMALL); phase = 2; break; }
                    FSEB_DSTEP(s1, sy); out[op++] = (u8)sy; ch_fill(c, win32);
                    if (ch_reload(c) == SRC_OVER) { FSEB_DSTEP(s2, sy); out[op++] = (u8)sy; ret = (u64)op; phase = 2; break; }
                    if (op > omax - 2) { ret = err(E_DST_TOO_SMALL); phase = 2; break; }
                    FSEB_DSTEP(s2, sy); out[op++] = (u8)sy; ch_fill(c, win32);
                    if (ch_reload(c) == SRC_OVER) { FSEB_DSTEP(s1, sy); out[op++] = (u8)sy; ret = (u64)op; phase = 2; break; }
                }
            }
        }
    }
    return __shfl_sync(FULL, ret, 0);
}

// FSE_decompressU16_usingDTable (fseU16.c:273-301).  cap / return value in symbols.
__device__ inline u64 warp_decode_u16(u16* out, u64 cap, const u8* cSrc, u64 cSize, const u32* dt, u64* win)
{
    unsigned const lane = lane_id();
    unsigned const tl = dt[0] & 0xFFFF;
    const u32* const cells = dt + 1;
    const u32* const win32 = reinterpret_cast<const u32*>(win);
    u64 op = 0;
    if (cSize < 1) return err(E_CORRUPT);               // the reference dereferences a NULL stream here (documented deviation)
    WSrc b; b.base = reinterpret_cast<u64>(cSrc); b.len = cSize; b.winBase = 0; b.win = win; b.at = 0; b.w = 0; b.used = 0;
    Chain c; c.hi = c.lo = c.vb = 0; c.noff = 0; c.at = 0; c.used = 0;
    u32 st = 0; u64 ret = 0; int phase = 0;             // 0 = stream phase, 1 = drain phase, 2 = done
    ws_slide(b, cSize >= 8 ? cSize - 8 : 0, win, true);
    if (lane == 0) {
        u64 const e = ch_open(c, b, win32);             // the reference ignores the verdict (:286) and goes on with bitsConsumed = 0 (+ padding)
        if (is_err(e)) {                                // missing end mark: nothing is skipped, the stream is read from its last bit
            u64 const endA = b.base + cSize, top4 = (endA + 3) & ~3ull;
            c.hi = c.lo = c.vb = 0; c.noff = (u32)(top4 - 4 - b.winBase);
            ch_fill(c, win32);
            u32 const skip = (u32)(8 * (top4 - endA));
            (void)ch_take(c, skip > 16 ? 16 : skip); if (skip > 16) (void)ch_take(c, skip - 16);
            ch_fill(c, win32);
            c.used = 0;                                 // bitstream.h:283 / :304: bitsConsumed stays 0 when the last byte is 0
            c.at = cSize >= 8 ? cSize - 8 : 0;
            if (cSize < 8) {                            // short container: the stream sits in its low bytes under (8 - len) zero bytes
                u64 W = ((u64)c.hi << 32) | c.lo; W >>= (8 - cSize) * 8;
                c.hi = (u32)(W >> 32); c.lo = (u32)W; c.vb = 64;
            }
        }
        st = ch_take(c, tl); ch_reload(c); ch_fill(c, win32);
    }
#define FSEB_U16_STEP() do { u32 const cell = cells[st]; out[op++] = (u16)(cell >> 20); \
                             st = (cell & 0xFFFF) + ch_take(c, (cell >> 16) & 0xF); ch_fill(c, win32); } while (0)
    for (;;) {
        phase = __shfl_sync(FULL, phase, 0);
        if (phase == 2) break;
        ch_slide(b, c, win);
        if (lane == 0) {
            if (phase == 0) {
                for (int it = 0; it < 24; it++) {       // <= 24 * 16 bits = 48 bytes per batch
                    if (!(ch_reload(c) < SRC_DONE && op < cap)) { phase = 1; break; }
                    FSEB_U16_STEP();
                }
            } else {
                if (!(c.at == 0 && c.used == 64)) { ret = err(E_CORRUPT); phase = 2; }
                else {
                    int it = 0;
                    for (; it < 64 && st && op < cap; it++) FSEB_U16_STEP();
                    if (!(st && op < cap)) { ret = st ? err(E_CORRUPT) : op; phase = 2; }
                }
            }
        }
    }
#undef FSEB_U16_STEP
    return __shfl_sync(FULL, ret, 0);
}
#undef FSEB_DSTEP

// Builds the DTable image of one block with one warp.  Returns 0 or an error (uniform across the warp).
template <bool WIDE>
__device__ inline u64 warp_build_dtable(DecWarp<WIDE>& w, unsigned msv, unsigned tl)
{
    unsigned const lane = lane_id();
    u32 const size = 1u << tl;
    if (msv > DecCfg<WIDE>::MSV) return err(E_MSV_TOO_LARGE);
    if (tl > DecCfg<WIDE>::TL) return err(E_TLOG_TOO_LARGE);
    unsigned fast = 1;
    if (lane == 0) {
        u32 acc = 0;
        for (u32 s = 0; s <= msv; s++) {
            w.cum[s] = (u16)acc;
            int const n = w.norm[s];
            if (n == -1) w.nextOf[s] = 1;
            else { if (n >= (int)(1u << (tl - 1))) fast = 0; w.nextOf[s] = (u16)n; if (n > 0) acc += (u32)n; }
        }
        w.cum[msv + 1] = (u16)acc;
        w.dt[0] = tl | (fast << 16);
    }
    __syncwarp();
    u32* const cells = w.dt + 1;
    bool const closed = warp_spread(w.norm, w.cum, msv, tl, [&](u32 cell, u32 sym) { cells[cell] = sym; });
    __syncwarp();
    if (!closed) return err(E_GENERIC);
    // second pass in cell order: the k-th cell (ascending) of symbol s gets x = norm[s] + k  (fse_decompress.c:117-124)
    for (u32 u0 = 0; u0 < size; u0 += 32) {
        u32 const u = u0 + lane;
        u32 const sym = cells[u];
        u32 const peers = __match_any_sync(FULL, sym);
        u32 const x = w.nextOf[sym] + __popc(peers & ((1u << lane) - 1));
        __syncwarp();
        if ((peers >> lane) == 1u) w.nextOf[sym] = (u16)(w.nextOf[sym] + __popc(peers));
        u32 const nb = tl - hibit(x);
        u32 const ns = ((x << nb) - size) & 0xFFFF;
        cells[u] = WIDE ? (ns | (nb << 16) | (sym << 20)) : (ns | (sym << 16) | (nb << 24));
        __syncwarp();
    }
    return 0;
}

// codec: 0 = FSE bytes, 2 = FSE U16 (sizes in the geometry are BYTES; U16 blocks hold blockSize/2 symbols)
template <bool WIDE>
__global__ void __launch_bounds__(THREADS)
fse_decode_kernel(BatchGeom g, u8* __restrict__ dst, const u8* __restrict__ cbuf, const u64* __restrict__ csizes,
                  u64* __restrict__ results, const u8* __restrict__ orig)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecWarp<WIDE>& w = reinterpret_cast<DecWarp<WIDE>*>(smem_raw)[threadIdx.x >> 5];
    unsigned const lane = lane_id();
    u32 const b = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (b >= g.nBlocks) return;
    u32 const n = block_len(g, b);
    u64 const cs = csizes[b];
    u8* const out = dst + (u64)b * g.blockSize;
    const u8* const c = cbuf + (u64)b * g.slot;

    // the harness' conventions for stored blocks (bench.c:393-402; U16 harness :247-262 only knows "raw")
    if (is_err(cs)) { if (lane == 0) results[b] = cs; return; }
    if (cs == 0 || (cs == 1 && !WIDE)) {
        if (orig) {
            const u8* const o = orig + (u64)b * g.blockSize;
            if (cs == 0) for (u32 i = lane; i < n; i += 32) out[i] = o[i];
            else { u8 const v = o[0]; for (u32 i = lane; i < n; i += 32) out[i] = v; }
        }
        if (lane == 0) results[b] = orig ? n : 0;
        return;
    }
    // header
    u64 h = 0; unsigned tl = 0, msv = DecCfg<WIDE>::MSV;
    if (lane == 0) {
        if (WIDE && cs < 2) h = err(E_SRC_WRONG);                                   // fseU16.c:317
        else h = d_read_ncount(w.norm, &msv, &tl, c, cs);
        if (!WIDE && !is_err(h) && tl > FSE_MAX_TLOG) h = err(E_TLOG_TOO_LARGE);    // fse_decompress.c:266
    }
    h = __shfl_sync(FULL, h, 0); tl = __shfl_sync(FULL, tl, 0); msv = __shfl_sync(FULL, msv, 0);
    if (is_err(h)) { if (lane == 0) results[b] = h; return; }
    __syncwarp();
    u64 const e = warp_build_dtable<WIDE>(w, msv, tl);
    if (is_err(e)) { if (lane == 0) results[b] = e; return; }
    __syncwarp();
    {
        u64 r;
        if (WIDE) { r = warp_decode_u16(reinterpret_cast<u16*>(out), n / 2, c + h, cs - h, w.dt, w.win); if (!is_err(r)) r *= 2; }
        else r = warp_decode_bytes(out, n, c + h, cs - h, w.dt, w.win);
        if (lane == 0) results[b] = r;
    }
}

// =================================================================================================
// encode
// =================================================================================================
template <bool WIDE> struct EncCfg;
template <> struct EncCfg<false> { static constexpr unsigned MSV = FSE_MAX_SV, CELLS = 1u << FSE_MAX_TLOG; typedef u8 sym_t; };
template <> struct EncCfg<true>  { static constexpr unsigned MSV = U16_MAX_SV, CELLS = 1u << FSE_MAX_TLOG; typedef u16 sym_t; };   // tableLog <= 12 as linked

template <bool WIDE>
struct EncWarp {
    u32   ct[1 + EncCfg<WIDE>::CELLS / 2 + 2 * (EncCfg<WIDE>::MSV + 1)];
    u32   count[EncCfg<WIDE>::MSV + 1];
    u32   start[EncCfg<WIDE>::MSV + 3];
    short norm[EncCfg<WIDE>::MSV + 1];
    u16   cum[EncCfg<WIDE>::MSV + 3];
    typename EncCfg<WIDE>::sym_t cellSym[EncCfg<WIDE>::CELLS];
    u32   cdfs[64], cdnb[64];     // deltaFindState / deltaNbBits of the 64 symbols of the current group (filled by the whole warp)
    u32   slot[64];               // (value | nbBits << 16) of up to 64 consecutive symbols, emission order
    u32   words[32];
};

// Appends the (value,nbBits) pairs in slot[0..cnt) to the stream.  Stream state: `carry` = pending low
// bits (carryBits < 32) that precede, `wpos` = index of the next 32-bit word of the word-aligned image
// starting at `base32` (global, 4-byte aligned).  Words beyond capWords are dropped (overflow is judged
// from the bit total, bitstream.h:246,258).
template <bool WIDE>
__device__ __forceinline__ void warp_emit(EncWarp<WIDE>& w, u32 cnt, u32& carry, u32& carryBits, u32& wpos,
                                          u32* base32, u32 capWords, u32 mis, u64& totalBits)
{
    unsigned const lane = lane_id();
    for (u32 base = 0; base < cnt; base += 32) {
        u32 const e = (base + lane < cnt) ? w.slot[base + lane] : 0u;
        u32 const nb = e >> 16, val = e & ((1u << nb) - 1);
        u32 incl = nb;                                             // inclusive scan of bit lengths
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 const t = __shfl_up_sync(FULL, incl, d); if (lane >= (unsigned)d) incl += t; }
        u32 const sum = __shfl_sync(FULL, incl, 31);
        u32 const off = carryBits + incl - nb;
        if (lane < 16) w.words[lane] = (lane == 0) ? carry : 0u;
        __syncwarp();
        if (nb) {
            u64 const v = (u64)val << (off & 31);
            atomicOr(&w.words[off >> 5], (u32)v);
            if ((u32)(v >> 32)) atomicOr(&w.words[(off >> 5) + 1], (u32)(v >> 32));
        }
        __syncwarp();
        u32 const tot = carryBits + sum;
        u32 const full = tot >> 5;
        if (lane < full && wpos + lane < capWords) {
            if (wpos + lane == 0 && mis) {                          // word 0 also covers `mis` bytes that belong to the header: leave them alone
                u8* const p8 = reinterpret_cast<u8*>(base32);
                for (u32 i = mis; i < 4; i++) p8[i] = (u8)(w.words[0] >> (8 * i));
            } else base32[wpos + lane] = w.words[lane];
        }
        carry = w.words[full]; carryBits = tot & 31; wpos += full; totalBits += sum;
        __syncwarp();
    }
}

template <bool WIDE>
__global__ void __launch_bounds__(THREADS)
fse_encode_kernel(BatchGeom g, u8* __restrict__ cbuf, u64* __restrict__ csizes, const u8* __restrict__ src,
                  unsigned msvReq, unsigned tlogReq)
{
    typedef typename EncCfg<WIDE>::sym_t sym_t;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EncWarp<WIDE>& w = reinterpret_cast<EncWarp<WIDE>*>(smem_raw)[threadIdx.x >> 5];
    unsigned const lane = lane_id();
    u32 const b = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (b >= g.nBlocks) return;
    u32 const nBytes = block_len(g, b);
    u32 const n = WIDE ? nBytes / 2 : nBytes;                        // symbols
    const sym_t* const s = reinterpret_cast<const sym_t*>(src + (u64)b * g.blockSize);
    u8* const d = cbuf + (u64)b * g.slot;
    u64 const cap = g.slot;
    unsigned const MSVMAX = EncCfg<WIDE>::MSV;
#define FSEB_DONE(v) do { if (lane == 0) csizes[b] = (v); return; } while (0)

    // ---- argument checks (fse_compress.c:645-649,691 ; fseU16.c:216-220) ----
    unsigned msv = msvReq, tl = tlogReq;
    if (!WIDE) {
        if (tl > FSE_MAX_TLOG) FSEB_DONE(err(E_TLOG_TOO_LARGE));
        if (tl == 0) FSEB_DONE(err(E_TLOG_TOO_LARGE));               // FSE_WKSP_SIZE_U32(0,..) shifts by -1: the reference (gcc/x86-64) reports tableLog_tooLarge
        if ((u64)14340 < (u64)1 + (1ull << (tl - 1)) + ((u64)msv + 1) * 2 + 1024) FSEB_DONE(err(E_TLOG_TOO_LARGE));   // :645 vs the stack workspace of :679-685
        if (n <= 1) FSEB_DONE(0);
        if (!msv) msv = FSE_MAX_SV;
        if (msv > FSE_MAX_SV) msv = FSE_MAX_SV;                      // HIST_count_wksp clamps (hist.c:171)
    } else {
        if (n <= 1) FSEB_DONE(n);
        if (!msv) msv = U16_MAX_SV;
        if (!tl) tl = U16_DEF_TLOG;
        if (msv > U16_MAX_SV) FSEB_DONE(err(E_MSV_TOO_LARGE));
        if (tl > U16_MAX_TLOG) FSEB_DONE(err(E_TLOG_TOO_LARGE));
    }
    // ---- histogram ----
    for (u32 i = lane; i <= MSVMAX; i += 32) w.count[i] = 0;
    __syncwarp();
    u32 over = 0;
    for (u32 i0 = 0; i0 < n; i0 += 32) {
        u32 const i = i0 + lane;
        u32 const v = (i < n) ? (u32)s[i] : 0xFFFFFFFFu;
        u32 const peers = __match_any_sync(FULL, v);
        if (i < n) {
            if (v > (WIDE ? msv : 255u)) over = 1;
            else if ((peers >> lane) == 1u) atomicAdd(&w.count[v], (u32)__popc(peers));
        }
    }
    __syncwarp();
    over = __any_sync(FULL, over);
    if (WIDE && over) FSEB_DONE(err(E_MSV_TOO_SMALL));               // fseU16.c:131
    u32 top = 0, best = 0;
    for (u32 i = lane; i <= (WIDE ? msv : 255u); i += 32) { u32 const c = w.count[i]; if (c) top = i; best = c > best ? c : best; }
    #pragma unroll
    for (int dlt = 16; dlt; dlt >>= 1) { top = max(top, __shfl_xor_sync(FULL, top, dlt)); best = max(best, __shfl_xor_sync(FULL, best, dlt)); }
    if (!WIDE && msv < 255 && top > msv) FSEB_DONE(err(E_MSV_TOO_SMALL));            // hist.c:128
    msv = top;
    if (best == n) FSEB_DONE(1);                                      // rle
    if (!WIDE) {
        if (best == 1) FSEB_DONE(0);
        if (best < (n >> 7)) FSEB_DONE(0);
    }
    // ---- normalise + header (one lane) ----
    tl = d_optimal_tablelog(tl, n, msv, 2);
    u64 hdr = 0;
    if (lane == 0) {
        hdr = d_normalize(w.norm, tl, w.count, n, msv);
        if (!is_err(hdr)) hdr = d_write_ncount(d, cap, w.norm, msv, tl);
    }
    hdr = __shfl_sync(FULL, hdr, 0);
    if (is_err(hdr)) FSEB_DONE(hdr);
    u32 const hSize = (u32)hdr;
    // ---- CTable (fse_compress.c:66-169) ----
    u32 const size = 1u << tl;
    if (lane == 0) {
        u32 acc = 0, st = 0;
        for (u32 q = 0; q <= msv; q++) {
            w.cum[q] = (u16)acc; w.start[q] = st;
            int const nn = w.norm[q];
            if (nn > 0) acc += (u32)nn;
            st += (u32)(nn == -1 ? 1 : nn);
        }
        w.cum[msv + 1] = (u16)acc;
        ((u16*)w.ct)[0] = (u16)tl; ((u16*)w.ct)[1] = (u16)msv;
    }
    __syncwarp();
    warp_spread(w.norm, w.cum, msv, tl, [&](u32 cell, u32 sym) { w.cellSym[cell] = (typename EncCfg<WIDE>::sym_t)sym; });
    __syncwarp();
    {   u16* const next = ((u16*)w.ct) + 2;
        for (u32 u0 = 0; u0 < size; u0 += 32) {                       // next[start[s] + k] = size + u for the k-th cell u of s (:125-128)
            u32 const u = u0 + lane;
            u32 const sym = w.cellSym[u];
            u32 const peers = __match_any_sync(FULL, sym);
            u32 const at = w.start[sym] + __popc(peers & ((1u << lane) - 1));
            next[at] = (u16)(size + u);
            __syncwarp();
            if ((peers >> lane) == 1u) w.start[sym] += __popc(peers);
            __syncwarp();
        }
        u32* const tt = w.ct + 1 + (size >> 1);
        if (lane == 0) {
            u32 total = 0;
            for (u32 q = 0; q <= msv; q++) {
                int const nn = w.norm[q];
                if (nn == 0) { tt[2 * q + 1] = ((tl + 1) << 16) - size; tt[2 * q] = 0; continue; }
                if (nn == -1 || nn == 1) { tt[2 * q + 1] = (tl << 16) - size; tt[2 * q] = total - 1; total++; }
                else {
                    u32 const maxOut = tl - hibit((u32)nn - 1);
                    tt[2 * q + 1] = (maxOut << 16) - ((u32)nn << maxOut);
                    tt[2 * q] = total - (u32)nn; total += (u32)nn;
                }
            }
        }
    }
    __syncwarp();
    // ---- encode (fse_compress.c:554-611 ; U16: fseU16.c:150-200) ----
    u64 const scap = cap - hSize;                                    // capacity seen by the stream writer
    bool const usable = scap > 8;
    if (!WIDE && (n <= 2 || !usable)) FSEB_DONE(0);
    u8* const sbase = d + hSize;
    u32 const mis = (u32)(reinterpret_cast<u64>(sbase) & 3);         // image words are aligned; the first one also holds `mis` header bytes
    u32* const base32 = reinterpret_cast<u32*>(sbase - mis);
    u32 carry = 0, carryBits = 8 * mis, wpos = 0;
    // the low `mis` bytes of image word 0 are header bytes; they are never rewritten, so their value is irrelevant (kept 0)
    u32 const capWords = (u32)((mis + (usable ? scap : 0)) / 4);
    u64 totalBits = 0;
    const u16* const next = ((const u16*)w.ct) + 2;
    const u32* const tt = w.ct + 1 + (size >> 1);
    u32 state = 0; bool seeded = false;
    if (WIDE) { state = size; seeded = true; }
    // symbols are consumed from the end; index parity picks the chain (lane 0: even / the only chain, lane 1: odd)
    for (u32 hi = n; hi > 0;) {
        u32 const cnt = hi >= 64 ? 64 : hi;
        u32 const lo = hi - cnt;
        // slot j <-> symbol index hi-1-j.  The whole warp first fetches the group's symbols (coalesced) and their
        // transforms, so that the chain lanes only touch shared memory and their loads do not depend on the state.
        for (u32 j = lane; j < cnt; j += 32) {
            u32 const sym = (u32)s[hi - 1 - j];
            w.cdfs[j] = tt[2 * sym]; w.cdnb[j] = tt[2 * sym + 1];
        }
        __syncwarp();
        {
            unsigned const chains = WIDE ? 1u : 2u;
            if (lane < chains) {
                u32 const j0 = WIDE ? 0u : ((((hi - 1) & 1u) == lane) ? 0u : 1u);      // first slot of this chain in the group
                for (u32 j = j0; j < cnt; j += chains) {
                    u32 const dfs = w.cdfs[j], dnb = w.cdnb[j];
                    if (!seeded) {                                   // FSE_initCState2 (fse.h:503-512): no output
                        u32 const nb0 = (dnb + (1u << 15)) >> 16;
                        u32 const v0 = (nb0 << 16) - dnb;
                        state = next[(v0 >> nb0) + dfs];
                        seeded = true;
                        w.slot[j] = 0;
                    } else {                                         // FSE_encodeSymbol (fse.h:514-521)
                        u32 const nb = (state + dnb) >> 16;
                        w.slot[j] = (state & ((1u << nb) - 1)) | (nb << 16);
                        state = next[(state >> nb) + dfs];
                    }
                }
            }
        }
        __syncwarp();
        warp_emit<WIDE>(w, cnt, carry, carryBits, wpos, base32, capWords, mis, totalBits);
        hi = lo;
    }
    // final states (fse_compress.c:608-609), end mark, close (bitstream.h:254-260)
    {   u32 const s0 = __shfl_sync(FULL, state, 0), s1 = __shfl_sync(FULL, state, 1);
        u32 k = 0;
        if (lane == 0) {
            if (!WIDE) w.slot[k++] = (s1 & (size - 1)) | (tl << 16);
            w.slot[k++] = (s0 & (size - 1)) | (tl << 16);
            w.slot[k++] = 1u | (1u << 16);
        }
        k = WIDE ? 2 : 3;
        __syncwarp();
        warp_emit<WIDE>(w, k, carry, carryBits, wpos, base32, capWords, mis, totalBits);
    }
    u64 streamBytes;
    if (!usable || (totalBits >> 3) >= scap - 8) streamBytes = 0;     // bitstream.h:190,258
    else {
        streamBytes = (totalBits + 7) >> 3;
        if (lane == 0 && carryBits > 8 * mis * (wpos == 0)) {         // flush the partial last word byte by byte
            u32 const nby = (carryBits + 7) / 8;
            u8* const p = reinterpret_cast<u8*>(base32 + wpos);
            for (u32 i = (wpos == 0 ? mis : 0); i < nby; i++) p[i] = (u8)(carry >> (8 * i));
        }
    }
    if (!WIDE) {
        if (streamBytes == 0) FSEB_DONE(0);                           // fse_compress.c:669
        u64 const totalOut = hSize + streamBytes;
        if (totalOut >= (u64)n - 1) FSEB_DONE(0);                     // :674
        FSEB_DONE(totalOut);
    } else {
        u64 const totalOut = hSize + streamBytes;
        if (totalOut >= (u64)(n - 1) * 2) FSEB_DONE(0);               // fseU16.c:248
        FSEB_DONE(totalOut);
    }
#undef FSEB_DONE
}

}  // namespace fsek

template <bool WIDE>
static cudaError_t launch_dec(const BatchGeom& g, void* dst, const void* cbuf, const u64* csizes, u64* results, const void* orig, cudaStream_t stream)
{
    if (g.nBlocks == 0) return cudaSuccess;
    size_t const smem = sizeof(fsek::DecWarp<WIDE>) * fsek::WARPS;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fsek::fse_decode_kernel<WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    unsigned const grid = (g.nBlocks + fsek::WARPS - 1) / fsek::WARPS;
    fsek::fse_decode_kernel<WIDE><<<grid, fsek::THREADS, smem, stream>>>(g, (u8*)dst, (const u8*)cbuf, csizes, results, (const u8*)orig);
    return cudaGetLastError();
}
template <bool WIDE>
static cudaError_t launch_enc(const BatchGeom& g, void* cbuf, u64* csizes, const void* src, unsigned msv, unsigned tlog, cudaStream_t stream)
{
    if (g.nBlocks == 0) return cudaSuccess;
    size_t const smem = sizeof(fsek::EncWarp<WIDE>) * fsek::WARPS;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fsek::fse_encode_kernel<WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    unsigned const grid = (g.nBlocks + fsek::WARPS - 1) / fsek::WARPS;
    fsek::fse_encode_kernel<WIDE><<<grid, fsek::THREADS, smem, stream>>>(g, (u8*)cbuf, csizes, (const u8*)src, msv, tlog);
    return cudaGetLastError();
}

cudaError_t launch_fse_decode(const BatchGeom& g, void* dst, const void* cbuf, const u64* csizes, u64* results, const void* orig, cudaStream_t s)
{ return launch_dec<false>(g, dst, cbuf, csizes, results, orig, s); }
cudaError_t launch_fseu16_decode(const BatchGeom& g, void* dst, const void* cbuf, const u64* csizes, u64* results, const void* orig, cudaStream_t s)
{ return launch_dec<true>(g, dst, cbuf, csizes, results, orig, s); }
cudaError_t launch_fse_encode(const BatchGeom& g, void* cbuf, u64* csizes, const void* src, unsigned msv, unsigned tlog, cudaStream_t s)
{ return launch_enc<false>(g, cbuf, csizes, src, msv, tlog, s); }
cudaError_t launch_fseu16_encode(const BatchGeom& g, void* cbuf, u64* csizes, const void* src, unsigned msv, unsigned tlog, cudaStream_t s)
{ return launch_enc<true>(g, cbuf, csizes, src, msv, tlog, s); }

}  // namespace fseb
