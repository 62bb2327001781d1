// huf_decode.cu -- batched Huff0 4-stream decode for sm_100a (HBM-bound integer path, no tensor cores).
//
// Replaces, for a whole batch of independent blocks, the CPU chain
//   HUF_decompress            lib/huf_decompress.c:1056-1081  (raw / RLE / decoder choice)
//   HUF_readStats             lib/entropy_common.c:154-215
//   HUF_readDTableX1          lib/huf_decompress.c:118-185
//   HUF_decompress4X1_usingDTable_internal_body  lib/huf_decompress.c:262-354
// Decoded bytes and error verdicts are those of the single-symbol (X1) decoder; X1 vs X2 is a CPU
// speed heuristic only (lib/huf_decompress.c:1029-1051) and both regenerate identical bytes.
//
// B200 mapping ("lane per stream, table per bank"):
//   * one CTA = 64 consecutive blocks = 256 streams = 256 threads; a warp holds the SAME stream
//     index of 32 DIFFERENT blocks, so the 32 lanes of a table look-up hit 32 different tables;
//   * the per-block decode table is a 9-bit first-level table of {nbBits, symbol} cells, interleaved
//     as u16 main[512][64]: the cell of block j sits in column j, i.e. the 32 lanes of a warp
//     (blocks 2*lane + parity) read 32 different banks -> conflict-free LDS for a random index;
//   * codes longer than 9 bits (rare: their probability is < 2^-9 each) take a canonical-code
//     fallback from <= 3 rank thresholds and the per-block list of symbols sorted by code length;
//   * each lane keeps a 64-bit left-aligned bit window in two registers, refilled 32 bits at a time
//     from a private 8-word shared-memory ring that the lane itself stages with aligned 16-byte
//     global loads (one load in flight ahead of use), walking its stream backwards;
//   * output: each lane packs 4 symbols per 32-bit store into its own quarter of the block.
#include "common.cuh"
#include "huf_dev.cuh"

namespace fseb {
namespace hufd {

constexpr int G = 64;             // blocks per CTA
constexpr int THREADS = 4 * G;    // one lane per stream
constexpr int MAIN_BITS = 9;
constexpr int MAIN_ROWS = 1 << MAIN_BITS;
constexpr int RING = 8;           // 32-bit words of stream look-ahead per lane
constexpr u32 NOERR = 0xFFFFFFFFu;

struct __align__(16) Smem {
    u16 main[MAIN_ROWS][G];       // 64 KB   first-level table, column = block
    u32 ring[RING][THREADS];      //  8 KB   per-lane stream words (column = thread)
    u8  sorted[G][256];           // 16 KB   symbols ordered by (weight asc, symbol asc) == code order
    u16 rankEnd[HUF_MAX_TLOG + 2][G];   // end (exclusive) of weight w's range in tableLog-bit index space
    u16 listStart[HUF_MAX_TLOG + 2][G]; // first position of weight w in sorted[]
    u8  tlog[G];
    u8  kind[G];                  // 0 = Huffman, 1 = raw copy, 2 = RLE, 3 = done/skip
    u32 status[G];                // NOERR or (stage<<8 | error code), smallest wins
    u32 hsize[G];                 // header bytes
    u8  weights[THREADS / 32][256];
    u32 rankStats[THREADS / 32][HUF_MAX_TLOG + 1];
    u16 rankRun[THREADS / 32][HUF_MAX_TLOG + 2];
};

// canonical-code look-up in tableLog-bit index space -> (nbBits | symbol << 8)
__device__ __forceinline__ u32 canon_lookup(const Smem& sm, int blk, u32 idx, u32 tl, u32 wStart)
{
    u32 w = wStart;
    while (idx >= sm.rankEnd[w][blk]) w++;
    u32 const first = (w == 1) ? 0u : sm.rankEnd[w - 1][blk];
    u32 const k = sm.listStart[w][blk] + ((idx - first) >> (w - 1));
    return (tl + 1 - w) | ((u32)sm.sorted[blk][k] << 8);
}

// Builds the tables of block `blk` (group-local index) with one warp.
__device__ void setup_block(Smem& sm, int blk, const u8* csrc, u64 csize, int warp)
{
    unsigned const lane = lane_id();
    u8* const weights = sm.weights[warp];
    u32 nbSym = 0, tl = 0; u64 h = 0;
    if (lane == 0) {
        h = d_huf_read_stats(weights, 256, sm.rankStats[warp], &nbSym, &tl, csrc, csize);
        if (!is_err(h) && tl > HUF_MAX_TLOG) h = err(E_TLOG_TOO_LARGE);      // huf_decompress.c:143
        if (!is_err(h) && h >= csize) h = err(E_SRC_WRONG);                   // huf_decompress.c:426
        if (!is_err(h)) {
            // rank ranges (huf_decompress.c:151-156): weight w covers 2^(w-1) cells per symbol, longest codes first
            u32 acc = 0, pos = 0;
            sm.rankEnd[0][blk] = 0; sm.listStart[0][blk] = 0;
            for (u32 w = 1; w <= tl; w++) {
                sm.listStart[w][blk] = (u16)pos; sm.rankRun[warp][w] = (u16)pos;
                pos += sm.rankStats[warp][w];
                acc += sm.rankStats[warp][w] << (w - 1);
                sm.rankEnd[w][blk] = (u16)(acc > 0xFFFF ? 0xFFFF : acc);
            }
            sm.rankEnd[tl][blk] = (u16)((1u << tl) > 0xFFFF ? 0xFFFF : (1u << tl));
            sm.rankEnd[tl + 1][blk] = 0xFFFF;
            sm.tlog[blk] = (u8)tl;
            sm.hsize[blk] = (u32)h;
        } else {
            atomicMin(&sm.status[blk], (u32)(0u << 8 | (u32)(0 - h)));
        }
    }
    h = __shfl_sync(0xFFFFFFFFu, h, 0);
    if (is_err(h)) return;
    nbSym = __shfl_sync(0xFFFFFFFFu, nbSym, 0);
    tl = __shfl_sync(0xFFFFFFFFu, tl, 0);
    __syncwarp();
    // sorted symbol list: stable by weight, then symbol order (huf_decompress.c:158-183 fills cells in that order)
    for (u32 base = 0; base < nbSym; base += 32) {
        u32 const s = base + lane;
        u32 const w = (s < nbSym) ? weights[s] : 0u;
        u32 const peers = __match_any_sync(0xFFFFFFFFu, w);
        if (w) {
            u32 const before = __popc(peers & ((1u << lane) - 1));
            u32 const slot = sm.rankRun[warp][w] + before;
            sm.sorted[blk][slot] = (u8)s;
        }
        __syncwarp();
        if (w && (peers >> lane) == 1u) sm.rankRun[warp][w] = (u16)(sm.rankRun[warp][w] + __popc(peers));   // highest lane of the group
        __syncwarp();
    }
    // first-level table: cell = nbBits | symbol<<8 ; nbBits == 0 marks "longer than 9 bits"
    {   u32 w = 1;
        for (u32 i = 0; i < MAIN_ROWS / 32; i++) {
            u32 const idx9 = lane * (MAIN_ROWS / 32) + i;
            u32 const idx = (tl >= MAIN_BITS) ? (idx9 << (tl - MAIN_BITS)) : (idx9 >> (MAIN_BITS - tl));
            while (idx >= sm.rankEnd[w][blk]) w++;
            u32 const first = (w == 1) ? 0u : sm.rankEnd[w - 1][blk];
            u32 const k = sm.listStart[w][blk] + ((idx - first) >> (w - 1));
            u32 const nb = tl + 1 - w;
            sm.main[idx9][blk] = (u16)(nb <= MAIN_BITS ? (nb | ((u32)sm.sorted[blk][k] << 8)) : 0u);
        }
    }
}

__global__ void __launch_bounds__(THREADS, 2)
huf_decode_kernel(BatchGeom g, u8* __restrict__ dst, const u8* __restrict__ cbuf, const u64* __restrict__ csizes,
                  u64* __restrict__ results, const u8* __restrict__ orig, u32 flags)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    u32 const blk0 = blockIdx.x * G;

    if (tid < G) { sm.status[tid] = NOERR; sm.kind[tid] = 3; sm.hsize[tid] = 0; sm.tlog[tid] = 0; }
    __syncthreads();

    // ---- classify blocks and build tables: warp w handles blocks w, w+8, ... of the group ----
    for (int j = warp; j < G; j += THREADS / 32) {
        u32 const b = blk0 + j;
        if (b >= g.nBlocks) continue;                                   // warp-uniform
        u64 const n = block_len(g, b);
        u64 const cs = csizes[b];
        int kind;
        if (flags & 1u) kind = 0;                                       // HUF_decompress4X1 semantics: always a Huffman block
        else if (is_err(cs)) { kind = 3; if (lane == 0) results[b] = cs; }   // propagated compressor error
        else if (cs == 0) { kind = orig ? 1 : 3; if (lane == 0) results[b] = orig ? n : 0; }   // stored raw by the harness (bench.c:393-397)
        else if (n == 0) { kind = 3; if (lane == 0) results[b] = err(E_DST_TOO_SMALL); }        // huf_decompress.c:1063
        else if (cs > n) { kind = 3; if (lane == 0) results[b] = err(E_CORRUPT); }              // :1064
        else if (cs == n) kind = 1;                                                                // :1065
        else if (cs == 1) kind = 2;                                                                // :1066
        else kind = 0;
        if (lane == 0) sm.kind[j] = (u8)kind;
        if (kind == 0) setup_block(sm, j, cbuf + (u64)b * g.slot, cs, warp);
    }
    __syncthreads();

    // ---- per-lane stream set-up: thread -> (block column, stream) ----
    int const col = 2 * lane + (warp >> 2);          // block within group; a warp sees 32 distinct banks of main[][]
    int const strm = warp & 3;
    u32 const b = blk0 + col;
    bool const live = (b < g.nBlocks) && sm.kind[col] == 0 && sm.status[col] == NOERR;
    u32 const n = live ? block_len(g, b) : 0;
    u32 const seg = (n + 3) / 4;
    u32 segLen = 0;                                  // symbols this lane must produce
    u8* outp = dst + (u64)b * g.blockSize + (u64)strm * seg;
    u32 hi = 0, lo = 0, nw = 0, r = 0, k = 0, q = 0;
    u64 chunkTop = 0, sBegin = 0;                    // address just above chunk 0 ; first byte of the stream
    uint4 pend = make_uint4(0, 0, 0, 0);
    u32 const tl = sm.tlog[col];

    auto load_chunk = [&](u32 qq) -> uint4 {         // aligned 16 bytes ending at chunkTop - 16*qq ; bytes below the stream start read as 0
        u64 const top = chunkTop - 16ull * qq;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (top > sBegin && 16ull * qq < chunkTop) {
            u64 const a = top - 16;
            v = __ldg(reinterpret_cast<const uint4*>(a));
            if (a < sBegin) {                        // zero the bytes that precede the stream
                u32 const z = (u32)(sBegin - a);     // 1..15
                u32* p = reinterpret_cast<u32*>(&v);
                #pragma unroll
                for (int i = 0; i < 4; i++) {
                    int const zb = (int)z - 4 * i;
                    if (zb >= 4) p[i] = 0; else if (zb > 0) p[i] &= 0xFFFFFFFFu << (8 * zb);
                }
            }
        }
        return v;
    };
    auto stage = [&](uint4 v, u32 qq) {              // chunk qq holds words 4qq..4qq+3 in descending address order
        u32 const s0 = (4 * qq) & (RING - 1);
        sm.ring[s0 + 0][tid] = v.w; sm.ring[s0 + 1][tid] = v.z; sm.ring[s0 + 2][tid] = v.y; sm.ring[s0 + 3][tid] = v.x;
    };

    if (live) {
        const u8* const cs0 = cbuf + (u64)b * g.slot;
        u64 const cs = csizes[b];
        const u8* const pay = cs0 + sm.hsize[col];
        u64 const psize = cs - sm.hsize[col];
        u32 code = 0;
        if (psize < 10) code = E_CORRUPT;                                            // huf_decompress.c:268
        else if (3 * seg > n) code = E_CORRUPT;                                      // dst too small for 4 segments (documented deviation)
        else {
            u32 const l1 = rd16(pay), l2 = rd16(pay + 2), l3 = rd16(pay + 4);
            if ((u64)l1 + l2 + l3 + 6 > psize) code = E_CORRUPT;                     // :302 length4 overflow
            else {
                u32 const l4 = (u32)(psize - 6 - l1 - l2 - l3);
                u32 const off = 6 + (strm > 0 ? l1 : 0) + (strm > 1 ? l2 : 0) + (strm > 2 ? l3 : 0);
                u32 const len = strm == 0 ? l1 : strm == 1 ? l2 : strm == 2 ? l3 : l4;
                if (len < 1) code = E_SRC_WRONG;                                     // BIT_initDStream, bitstream.h:274
                else {
                    u8 const last = pay[off + len - 1];
                    if (last == 0) code = (len >= 8) ? E_GENERIC : E_CORRUPT;        // :282-284 / :303-306
                    else {
                        sBegin = (u64)(pay + off);
                        u64 const e = sBegin + len;                                  // one past the last byte
                        chunkTop = ((e - 1) & ~15ull) + 16;
                        u32 const c0 = (u32)(8 * (chunkTop - e)) + (8 - hibit(last));   // garbage above the stream + zero padding + end mark
                        segLen = (strm < 3) ? seg : n - 3 * seg;
                        stage(load_chunk(0), 0); stage(load_chunk(1), 1);
                        q = 2; pend = load_chunk(2);
                        k = c0 >> 5; r = c0 & 31;
                        u32 const w0 = sm.ring[k & (RING - 1)][tid], w1 = sm.ring[(k + 1) & (RING - 1)][tid];
                        nw = sm.ring[(k + 2) & (RING - 1)][tid];
                        hi = __funnelshift_l(w1, w0, r); lo = w1 << r;
                    }
                }
            }
        }
        if (code) atomicMin(&sm.status[col], (u32)((1u + strm) << 8 | code));
    }
    __syncthreads();                                  // init verdicts of all four streams are in
    bool const go = live && sm.status[col] == NOERR;
    if (!go) segLen = 0;

    // ---- decode ----
    const u16* const tab = &sm.main[0][col];
    u32 const wLong = 1;                              // long codes start at weight 1
    u32 const longShift = 32 - tl;
    auto refill = [&]() {                             // a 32-bit boundary was crossed: merge the prefetched word
        r &= 31; k++;
        hi |= __funnelshift_l(nw, 0, r); lo = nw << r;
        nw = sm.ring[(k + 2) & (RING - 1)][tid];
    };
    auto decode1 = [&]() -> u32 {                     // returns nbBits | symbol << 8 and advances the window
        u32 e = tab[(hi >> (32 - MAIN_BITS)) * G];
        if (__builtin_expect((e & 0xFF) == 0, 0)) e = canon_lookup(sm, col, hi >> longShift, tl, wLong);
        hi = __funnelshift_l(lo, hi, e); lo = __funnelshift_l(0, lo, e);
        r += e & 0xFF;
        return e;
    };
    auto top_up = [&]() {                             // keep the ring ahead of the consumer (at most one chunk per 8 symbols)
        if (4 * q <= k + 7) { stage(pend, q); q++; pend = load_chunk(q); }
    };

    u32 pos = 0;
    bool const aligned = ((reinterpret_cast<u64>(outp) & 7) == 0);   // 8-byte stores
    if (aligned) {
        u32 const nIter = segLen >> 3;
        for (u32 it = 0; it < nIter; it++) {
            top_up();
            u32 o0 = 0, o1 = 0, e;
            e = decode1(); o0 = __byte_perm(o0, e, 0x3215);
            e = decode1(); o0 = __byte_perm(o0, e, 0x3250);
            if (r >= 32) refill();
            e = decode1(); o0 = __byte_perm(o0, e, 0x3510);
            e = decode1(); o0 = __byte_perm(o0, e, 0x5210);
            if (r >= 32) refill();
            e = decode1(); o1 = __byte_perm(o1, e, 0x3215);
            e = decode1(); o1 = __byte_perm(o1, e, 0x3250);
            if (r >= 32) refill();
            e = decode1(); o1 = __byte_perm(o1, e, 0x3510);
            e = decode1(); o1 = __byte_perm(o1, e, 0x5210);
            if (r >= 32) refill();
            *reinterpret_cast<uint2*>(outp + pos) = make_uint2(o0, o1);
            pos += 8;
        }
    }
    while (pos < segLen) {                            // ragged tails and unaligned segments: one symbol at a time
        if ((pos & 7) == 0) top_up();
        u32 const e = decode1();
        if (r >= 32) refill();
        outp[pos++] = (u8)(e >> 8);
    }

    // ---- verdict: every stream must be consumed exactly (huf_decompress.c:348-349) ----
    if (go) {
        u64 const consumed = 32ull * k + r;
        u64 const expect = 8ull * (chunkTop - sBegin);
        if (consumed != expect) atomicMin(&sm.status[col], (u32)(5u << 8 | E_CORRUPT));
    }
    __syncthreads();
    if (tid < G) {
        u32 const bb = blk0 + tid;
        if (bb < g.nBlocks && sm.kind[tid] == 0) {
            u32 const st = sm.status[tid];
            results[bb] = (st == NOERR) ? (u64)block_len(g, bb) : err(st & 0xFF);
        }
    }
    // ---- raw / RLE blocks (huf_decompress.c:1065-1066; the harness' own 0-size convention, bench.c:393-402) ----
    for (int j = 0; j < G; j++) {
        int const kd = sm.kind[j];
        if (kd != 1 && kd != 2) continue;
        u32 const bb = blk0 + j;
        u32 const nn = block_len(g, bb);
        u64 const cs = csizes[bb];
        u8* const o = dst + (u64)bb * g.blockSize;
        if (kd == 2) { u8 const v = cbuf[(u64)bb * g.slot]; for (u32 i = tid; i < nn; i += THREADS) o[i] = v; }
        else {
            const u8* const s = (cs == 0) ? orig + (u64)bb * g.blockSize : cbuf + (u64)bb * g.slot;
            for (u32 i = tid; i < nn; i += THREADS) o[i] = s[i];
        }
        if (tid == 0) results[bb] = nn;
    }
}

}  // namespace hufd

cudaError_t launch_huf_decode(const BatchGeom& g, void* dst, const void* cbuf, const u64* csizes, u64* results,
                              const void* orig, cudaStream_t stream, u32 flags)
{
    static bool configured = false;
    size_t const smem = sizeof(hufd::Smem);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(hufd::huf_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (g.nBlocks == 0) return cudaSuccess;
    unsigned const grid = (g.nBlocks + hufd::G - 1) / hufd::G;
    hufd::huf_decode_kernel<<<grid, hufd::THREADS, smem, stream>>>(g, (u8*)dst, (const u8*)cbuf, csizes, results, (const u8*)orig, flags);
    return cudaGetLastError();
}

}  // namespace fseb
