// huf_decode.cu -- batched Huff0 4-stream decode for sm_100a (HBM-bound integer path, no tensor cores).
//
// Replaces, for a whole batch of independent blocks, the CPU chain
//   HUF_decompress            lib/huf_decompress.c:1056-1081  (raw / RLE / decoder choice)
//   HUF_readStats             lib/entropy_common.c:154-215
//   HUF_readDTableX1          lib/huf_decompress.c:118-185
//   HUF_decompress4X1_usingDTable_internal_body  lib/huf_decompress.c:262-354
// Decoded bytes are those of both CPU decoders (they agree on every valid stream); the error verdict produced here is
// the single-symbol (X1) decoder's.  Where the reference would have run its double-symbol decoder
// (HUF_selectDecoder, lib/huf_decompress.c:1029-1051) and this kernel rejects a stream, huf_x2_fixup.cu re-runs the
// block under the exact X2 end-of-stream rules (a stream accepted here is accepted by X2 with the same bytes).
//
// B200 mapping ("lane per stream, table column per bank"), round 2:
//   * one CTA = 64 consecutive blocks = 256 streams = 256 threads; a warp holds the SAME stream index of 32
//     DIFFERENT blocks (the even or the odd columns), so the 32 look-ups of a warp instruction hit 32 tables;
//   * ONE unified u16 table per block, column-interleaved tbl[row][64]: rows [0, CUT) hold full-resolution cells
//     (indexed by the top tableLog bits of the window) for the few windows that start a code longer than M bits
//     (Huff0 numbers the longest codes from 0, so these are exactly the windows below a per-block threshold),
//     rows >= CUT hold the M-bit first-level table minus its never-used head.  The row of window value x is
//         min(x, D + (x >> (tableLog - M))),   D = CUT - (CUT >> (tableLog - M))
//     -- one IMNMX instead of a compare/select, and the split M is chosen PER BLOCK to minimise the rows
//     (P14: M = 7, 283 rows; the round-1 layout needed 704).  372 rows of 128 B + an 8 KB stream ring + 1 KB of
//     per-block facts = 55.75 KB per CTA: FOUR CTAs (1024 lanes = 256 blocks) per SM instead of two, so the 221
//     blocks per SM of a 1 GiB batch decode in one round with 28 warps resident;
//   * bit window: the lane keeps three raw 32-bit stream words w0..w2 and a bit offset r < 32; the 64-bit window is
//     recomputed from (w0,w1,w2) by two funnel shifts after every PAIR of symbols, and when r crosses 32 the words
//     rotate (predicated moves) and the next word is fetched from the lane's private shared-memory ring.  No
//     running 64-bit shift, no merge, 2 funnel shifts + 1 test per pair instead of round 1's 14-instruction refill;
//   * stream ring: 8 words per lane, word-interleaved ring[slot][tid] (bank = lane: conflict-free for any 32
//     cursors).  The lane tops it up itself with one 16-byte global load per 4 words, issued one top-up ahead into
//     registers (the load is in flight across ~16 symbols of work, nothing waits on it);
//   * output: 32 symbols are packed into one 32-byte store per lane (STG.256: a whole sector per store);
//   * blocks whose best table exceeds the row budget ("hard", e.g. near-flat 256-symbol alphabets), unaligned
//     segments and ragged tails take a per-symbol loop (canonical-code search for hard blocks).
#include "common.cuh"
#include "huf_dev.cuh"
#include "launch_util.cuh"
#include <cstdlib>

namespace fseb {
namespace hufd {

constexpr int G = 64;             // block columns per CTA
constexpr int THREADS = 4 * G;    // one lane per stream
constexpr int NWARPS = THREADS / 32;
constexpr int RW = 8;             // ring words per lane
constexpr u32 RING_BYTES = RW * THREADS * 4;
constexpr u32 NOERR = 0xFFFFFFFFu;
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr u32 MIN_ROWS = 160;     // a hard block parks its canonical-code arrays in 156 rows of its own column
constexpr u32 PARK_RANKEND = 128, PARK_LISTSTART = 142;

struct __align__(16) Facts {      // per-CTA facts about its 64 block columns (1.25 KB)
    u32 bid[G];                   // batch index of the block in this column (NOBLOCK = empty column)
    u32 status[G];                // NOERR or (stage<<8 | error code), smallest wins
    u32 hsize[G];                 // tree-header bytes
    u16 dOff[G];                  // D of the row formula
    u16 cut[G];
    u8  tlog[G];
    u8  mbits[G];                 // M; == tlog for a main-only table
    u8  kind[G];                  // 0 = Huffman, 1 = raw copy, 2 = RLE, 3 = done/skip
    u8  hard[G];
};
struct BuildScratch {             // lives in the (not yet used) stream ring while tables are built
    u8  weights[NWARPS][256];
    u8  sorted[NWARPS][256];      // symbols in code order: weight ascending, symbol ascending
    u32 rankStats[NWARPS][HUF_MAX_TLOG + 1];
    u16 rankRun[NWARPS][HUF_MAX_TLOG + 2];
    u16 rankEnd[NWARPS][HUF_MAX_TLOG + 2];     // end (exclusive) of weight w's range in tableLog-bit index space
    u16 listStart[NWARPS][HUF_MAX_TLOG + 2];   // first position of weight w in the sorted symbol list
};
static_assert(sizeof(BuildScratch) <= RING_BYTES, "build scratch must fit in the ring");
static_assert(sizeof(Facts) == 1280, "Facts layout");
constexpr u32 NOBLOCK = 0xFFFFFFFFu;

__host__ __device__ constexpr u32 smem_bytes(u32 rows) { return rows * (G * 2) + RING_BYTES + (u32)sizeof(Facts); }

__device__ __forceinline__ u32 lds_u16(u32 addr) { u16 v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr)); return v; }
__device__ __forceinline__ u32 lds_u32(u32 addr) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts_u32(u32 addr, u32 v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }

// canonical-code look-up in tableLog-bit index space -> (nbBits | symbol << 8)
__device__ __forceinline__ u32 canon_cell(const u16* rankEnd, const u16* listStart, const u8* sorted, u32 idx, u32 tl)
{
    u32 w = 1;
    while (idx >= rankEnd[w]) w++;
    u32 const first = (w == 1) ? 0u : rankEnd[w - 1];
    u32 const k = listStart[w] + ((idx - first) >> (w - 1));
    return (tl + 1 - w) | ((u32)sorted[k] << 8);
}

// HUF_readStats (lib/entropy_common.c:158-215) with the whole warp: same verdicts as d_huf_read_stats.  The raw 4-bit form and
// the per-weight statistics are spread over the lanes (they were 7 % of this kernel's instructions on one lane); the
// FSE-compressed form (a serial 2-state decode of <= 255 weights) stays on lane 0.  All lanes return the same value.
__device__ u64 warp_huf_read_stats(u8* weights, u32* rankStats, u32* nbSymPtr, u32* tlPtr, const u8* in, u64 srcSize, unsigned lane)
{
    if (!srcSize) return err(E_SRC_WRONG);
    u64 iSize = in[0], oSize;
    if (iSize >= 128) {                                   // raw 4-bit weights (:167-177)
        oSize = iSize - 127;
        iSize = (oSize + 1) / 2;
        if (iSize + 1 > srcSize) return err(E_SRC_WRONG);
        if (oSize >= 256) return err(E_CORRUPT);
        for (u32 k = lane; k < (u32)iSize; k += 32) { u8 const v = in[1 + k]; weights[2 * k] = v >> 4; weights[2 * k + 1] = v & 15; }
    } else {                                              // FSE-compressed weights (:178-183)
        if (iSize + 1 > srcSize) return err(E_SRC_WRONG);
        oSize = 0;
        if (lane == 0) oSize = d_huf_read_weights_fse(weights, 256, in, iSize);
        oSize = __shfl_sync(FULL, oSize, 0);
        if (is_err(oSize)) return oSize;
    }
    if (lane <= HUF_MAX_TLOG) rankStats[lane] = 0;
    __syncwarp();
    u32 total = 0; bool bad = false;
    for (u32 n = lane; n < (u32)oSize; n += 32) {
        u32 const w = weights[n];
        if (w >= HUF_MAX_TLOG) bad = true;
        else { atomicAdd(&rankStats[w], 1u); total += (1u << w) >> 1; }
    }
    if (__any_sync(FULL, bad)) return err(E_CORRUPT);
    #pragma unroll
    for (int d = 16; d; d >>= 1) total += __shfl_xor_sync(FULL, total, d);
    if (total == 0) return err(E_CORRUPT);
    u32 const tl = hibit(total) + 1;
    if (tl > HUF_MAX_TLOG) return err(E_CORRUPT);
    u32 const rest = (1u << tl) - total;
    u32 const lastW = hibit(rest) + 1;
    if ((1u << hibit(rest)) != rest) return err(E_CORRUPT);
    __syncwarp();
    if (lane == 0) { weights[oSize] = (u8)lastW; rankStats[lastW]++; }
    __syncwarp();
    if ((rankStats[1] < 2) || (rankStats[1] & 1)) return err(E_CORRUPT);
    *tlPtr = tl;
    *nbSymPtr = (u32)(oSize + 1);
    return iSize + 1;
}

// Builds the unified table of block column `blk` with one warp.
__device__ void setup_block(u16* tbl, Facts& fx, BuildScratch& bs, u32 rows, int blk, const u8* csrc, u64 csize, int warp)
{
    unsigned const lane = lane_id();
    u8* const weights = bs.weights[warp];
    u8* const sorted = bs.sorted[warp];
    u16* const rankEnd = bs.rankEnd[warp];
    u16* const listStart = bs.listStart[warp];
    u32 nbSym = 0, tl = 0, M = 0, cut = 0, nRows = 0;
    u64 h = warp_huf_read_stats(weights, bs.rankStats[warp], &nbSym, &tl, csrc, csize, lane);
    if (lane == 0) {
        if (!is_err(h) && tl > HUF_MAX_TLOG) h = err(E_TLOG_TOO_LARGE);      // huf_decompress.c:143
        if (!is_err(h) && h >= csize) h = err(E_SRC_WRONG);                   // huf_decompress.c:426
        if (!is_err(h)) {
            // rank ranges (huf_decompress.c:151-156): weight w covers 2^(w-1) cells per symbol, longest codes first
            u32 acc = 0, pos = 0;
            rankEnd[0] = 0; listStart[0] = 0;
            for (u32 w = 1; w <= tl; w++) {
                listStart[w] = (u16)pos; bs.rankRun[warp][w] = (u16)pos;
                pos += bs.rankStats[warp][w];
                acc += bs.rankStats[warp][w] << (w - 1);
                rankEnd[w] = (u16)(acc > 0xFFFF ? 0xFFFF : acc);
            }
            rankEnd[tl] = (u16)(1u << tl);
            rankEnd[tl + 1] = 0xFFFF;
            // split choice: first level of m bits + full resolution below CUT; rows = CUT + 2^m - CUT / 2^(tl-m)
            M = tl; cut = 0; nRows = 1u << tl;                                 // main-only
            for (u32 m = (tl > 10 ? 10 : tl - 1); m >= 4 && m < tl; m--) {
                u32 const g = 1u << (tl - m);
                u32 const T = rankEnd[tl - m];                                 // windows that start a code longer than m bits
                u32 const c = (T + g - 1) & ~(g - 1);
                u32 const r = c + (1u << m) - (c >> (tl - m));
                if (r < nRows) { nRows = r; M = m; cut = c; }
            }
            fx.hard[blk] = (u8)(nRows > rows);
            fx.tlog[blk] = (u8)tl;
            fx.mbits[blk] = (u8)M;
            fx.cut[blk] = (u16)cut;
            fx.dOff[blk] = (u16)(M < tl ? cut - (cut >> (tl - M)) : 0);
            fx.hsize[blk] = (u32)h;
        } else {
            atomicMin(&fx.status[blk], (u32)(0u << 8 | (u32)(0 - h)));
        }
    }
    h = __shfl_sync(FULL, h, 0);
    if (is_err(h)) return;
    M = __shfl_sync(FULL, M, 0); cut = __shfl_sync(FULL, cut, 0); nRows = __shfl_sync(FULL, nRows, 0);
    __syncwarp();
    // sorted symbol list: stable by weight, then symbol order (huf_decompress.c:158-183 fills cells in that order)
    for (u32 base = 0; base < nbSym; base += 32) {
        u32 const s = base + lane;
        u32 const w = (s < nbSym) ? weights[s] : 0u;
        u32 const peers = __match_any_sync(FULL, w);
        if (w) sorted[bs.rankRun[warp][w] + __popc(peers & ((1u << lane) - 1))] = (u8)s;
        __syncwarp();
        if (w && (peers >> lane) == 1u) bs.rankRun[warp][w] = (u16)(bs.rankRun[warp][w] + __popc(peers));   // highest lane of the group
        __syncwarp();
    }
    u16* const col = tbl + blk;
    if (nRows <= rows) {
        u32 const D = (M < tl) ? cut - (cut >> (tl - M)) : 0u;
        for (u32 row = lane; row < nRows; row += 32) {
            u32 const x = (row < cut || M == tl) ? row : ((row - D) << (tl - M));
            col[row * G] = (u16)canon_cell(rankEnd, listStart, sorted, x, tl);
        }
    } else {   // hard block: park the code-ordered symbol list (2 per cell) and the rank arrays in the column
        for (u32 r = lane; r < 128; r += 32) col[r * G] = (u16)(sorted[2 * r] | (sorted[2 * r + 1] << 8));
        if (lane < HUF_MAX_TLOG + 2) { col[(PARK_RANKEND + lane) * G] = rankEnd[lane]; col[(PARK_LISTSTART + lane) * G] = listStart[lane]; }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(THREADS, 4)
huf_decode_kernel(BatchGeom g, u8* __restrict__ dst, const u8* __restrict__ cbuf, const u64* __restrict__ csizes,
                  u64* __restrict__ results, const u8* __restrict__ orig, u32 flags, u32 gEff, u32 rows,
                  const u32* __restrict__ list, const u32* __restrict__ listCount, u32* __restrict__ deferList, u32* __restrict__ deferCount)
{
    // Two passes share this kernel.  Pass A (list == nullptr) walks the batch in order with the four-CTAs-per-SM table budget; a
    // block whose smallest table does not fit that budget is not decoded but appended to deferList.  Pass B (deferList == nullptr)
    // walks that list with a larger budget (fewer CTAs per SM); what does not fit even there takes the per-symbol path.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // The ring comes first: its address is then (compile-time base) + an offset the hot loop builds with one OR.
    unsigned char* const ringRaw = smem_raw;
    BuildScratch& bs = *reinterpret_cast<BuildScratch*>(ringRaw);
    Facts& fx = *reinterpret_cast<Facts*>(ringRaw + RING_BYTES);
    u16* const tbl = reinterpret_cast<u16*>(smem_raw + RING_BYTES + sizeof(Facts));      // [rows][G]
    int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    u32 const blk0 = blockIdx.x * gEff;
    u32 const nWork = list ? *listCount : g.nBlocks;
    if (blk0 >= nWork) return;                                              // pass B is launched for the worst case: most CTAs find nothing
    u32 const blkEnd = min(blk0 + gEff, nWork);                             // this CTA owns work items [blk0, blkEnd)

    if (tid < G) { fx.bid[tid] = (blk0 + tid < blkEnd) ? (list ? list[blk0 + tid] : blk0 + tid) : NOBLOCK; fx.status[tid] = NOERR; fx.kind[tid] = 3; fx.hsize[tid] = 0; fx.tlog[tid] = 0; fx.mbits[tid] = 0; fx.cut[tid] = 0; fx.dOff[tid] = 0; fx.hard[tid] = 0; }
    __syncthreads();

    // ---- classify blocks and build tables: warp w handles columns w, w+8, ... ----
    for (int j = warp; j < G; j += NWARPS) {
        u32 const b = fx.bid[j];
        if (b == NOBLOCK) continue;                                     // warp-uniform
        u64 const n = block_len(g, b);
        u64 const cs = csizes[b];
        int kind;
        if (flags & 1u) kind = 0;                                       // HUF_decompress4X1 / 4X2 semantics: always a Huffman block
        else if (is_err(cs)) { kind = 3; if (lane == 0) results[b] = cs; }   // propagated compressor error
        else if (cs == 0) { kind = orig ? 1 : 3; if (lane == 0) results[b] = orig ? n : 0; }   // stored raw by the harness (bench.c:393-397)
        else if (n == 0) { kind = 3; if (lane == 0) results[b] = err(E_DST_TOO_SMALL); }        // huf_decompress.c:1063
        else if (cs > n) { kind = 3; if (lane == 0) results[b] = err(E_CORRUPT); }              // :1064
        else if (cs == n) kind = 1;                                                                // :1065
        else if (cs == 1) kind = 2;                                                                // :1066
        else kind = 0;
        if (lane == 0) fx.kind[j] = (u8)kind;
        if (kind == 0) {
            setup_block(tbl, fx, bs, rows, j, cbuf + (u64)b * g.slot, cs, warp);
            if (deferList && fx.hard[j] && fx.status[j] == NOERR) {     // pass A: hand the block to the pass with the larger table budget
                if (lane == 0) { deferList[atomicAdd(deferCount, 1u)] = b; fx.kind[j] = 3; }
            }
        }
    }
    __syncthreads();                                  // tables and facts complete; the build scratch (ring) is free now

    // ---- per-lane stream set-up: thread -> (block column, stream) ----
    int const col = 2 * lane + (warp >> 2);          // a warp sees the 32 even (or odd) columns = 32 distinct banks
    int const strm = warp & 3;
    u32 const b = fx.bid[col];
    bool const live = (b != NOBLOCK) && fx.kind[col] == 0 && fx.status[col] == NOERR;
    u32 const n = live ? block_len(g, b) : 0;
    u32 const seg = (n + 3) / 4;
    u32 segLen = 0;                                  // symbols this lane must produce
    u8* outp = dst + (u64)b * g.blockSize + (u64)strm * seg;
    u64 chunkTop = 0;                                // address just above chunk 0
    u32 expectBits = 0;                              // stream bits between chunkTop and the first byte of the stream
    u32 c0 = 0;                                      // bits to skip at the top of chunk 0: garbage above the stream + zero padding + end mark
    u32 const tl = fx.tlog[col];

    if (live) {
        const u8* const cs0 = cbuf + (u64)b * g.slot;
        u64 const cs = csizes[b];
        const u8* const pay = cs0 + fx.hsize[col];
        u64 const psize = cs - fx.hsize[col];
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
                        u64 const sBegin = (u64)(pay + off);
                        u64 const e = sBegin + len;                                  // one past the last byte
                        chunkTop = ((e - 1) & ~31ull) + 32;
                        c0 = (u32)(8 * (chunkTop - e)) + (8 - hibit(last));
                        expectBits = (u32)(8 * (chunkTop - sBegin));
                        segLen = (strm < 3) ? seg : n - 3 * seg;
                    }
                }
            }
        }
        if (code) atomicMin(&fx.status[col], (u32)((1u + strm) << 8 | code));
    }
    __syncthreads();                                  // init verdicts of all four streams are in
    bool const go = live && fx.status[col] == NOERR;
    if (!go) segLen = 0;

    // ---- stream feeder ----
    // Stream word j (j = 0 is the word just below chunkTop) is the little-endian u32 at chunkTop - 4(j+1); chunk q = words
    // 4q..4q+3 = one aligned 16 bytes.  Below the first byte of the stream the feeder delivers whatever lies there (the
    // previous stream, the jump table, the header): the reference pads with zeros instead, but those bits can only be
    // CONSUMED by a stream that is already over-read -- a code that straddles the stream start decodes, under any padding,
    // to a length beyond the bits that are left (prefix property) -- so the verdict "consumed == stream bits" is the same,
    // and a rejected block's bytes are not part of the contract.  That keeps the hot loop to ONE unconditional 16-byte load:
    // with a second, conditional load path inlined next to it, ptxas shared a scoreboard slot between the two and every
    // 16 symbols the window shift waited a DRAM round trip for a load it does not depend on (ncu round 2: 13% of samples).
    // Chunks that would lie below the compressed buffer itself re-read its first 16 bytes (address clamp, never dereferenced
    // out of bounds).
    u32 const ringLane = (u32)__cvta_generic_to_shared(ringRaw) + tid * 4;      // + slot * (THREADS*4)
    // Global loads are 32 bytes (one whole sector, LDG.256) = a PAIR of chunks: the lane's scattered 16-byte loads fetched
    // every sector twice (nothing survives in an L1 that shares 228 KB with 223 KB of shared memory) and cost twice the LSU
    // wavefronts.  The pair waits in eight registers; its two chunks enter the ring one after the other.
    u32 const pLimit = go ? (u32)((chunkTop - (reinterpret_cast<u64>(cbuf) & ~31ull)) >> 5) - 1u : 0u;   // last pair at or above the buffer start
    u32 m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0, m5 = 0, m6 = 0, m7 = 0;        // pending pair, memory order (m7 = highest address = first consumed)
    auto load_pair = [&](u32 pr) {
        u64 const a = chunkTop - 32ull * (min(pr, pLimit) + 1);
        asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(m0), "=r"(m1), "=r"(m2), "=r"(m3), "=r"(m4), "=r"(m5), "=r"(m6), "=r"(m7) : "l"(a));
        // ... and the L2 is asked for the sector two pairs further down the stream: the demand load above was itself
        // announced that way, so it is an L2 hit by now -- the 16 symbols until its data is needed cover an L2 round trip,
        // they did not cover a DRAM one (ncu: the store of the pending chunk waited on it for 16% of all samples).
        asm volatile("prefetch.global.L2 [%0];" :: "l"(chunkTop - 32ull * (min(pr + 2, pLimit) + 1)));
    };

    u32 w0 = 0, w1 = 0, w2 = 0, r = 0;               // raw stream words: the window starts at bit r of w0 and spans w0..w2
    u32 kBase = 0;                                    // absolute index of the next stream word to fetch = kBase + (r >> 5): `r` is never reduced
    u32 q = 0;                                        // next chunk to enter the ring (chunks 0..q-1 are consumed or in it); chunk q -> slots (4q..4q+3) mod 8
    auto store_next = [&]() {                         // chunk q out of the pending pair; an odd q uses the pair up: fetch the next one
        u32 const a = ringLane + (q & 1) * (4 * THREADS * 4);
        bool const lowHalf = (q & 1) != 0;
        sts_u32(a, lowHalf ? m3 : m7); sts_u32(a + THREADS * 4, lowHalf ? m2 : m6);
        sts_u32(a + 2 * THREADS * 4, lowHalf ? m1 : m5); sts_u32(a + 3 * THREADS * 4, lowHalf ? m0 : m4);
        q++;
        if (!(q & 1)) load_pair(q >> 1);
    };
    // Words written but not yet fetched: U = 4q - (absolute index of the next word to fetch).  The ring is topped up at two
    // kinds of check, 8 symbols (at most 3 fetched words) apart: the group-start check stores the next chunk when U <= 4, the
    // mid-group check only when U <= 3.  Induction: U >= 5 after a group-start check, hence >= 2 at the mid-group one and
    // >= 4 after it, hence >= 1 at the next group start -- U stays in [1, 8] at every check, so it is recoverable from the
    // slot numbers alone, and the ring never runs dry whatever the code lengths.  On typical data only the group-start
    // check stores, so a load has 16 symbols of work to land before the next store of this WARP touches its registers (the
    // scoreboard is per warp, not per lane: a store every 8 symbols exposed the latency).
    auto unread = [&]() -> u32 { return 4 * q - kBase - (r >> 5); };
    auto top_up = [&](u32 threshold) { if (unread() <= threshold) store_next(); };

    if (go) {                                         // the first pair fills the ring, the window words come out of it
        u32 const k0 = c0 >> 5;                       // first word of the window: 0..8
        r = c0 & 31;
        q = 2 * (k0 >> 3);
        load_pair(q >> 1);
        store_next(); store_next();                   // ring full: words 4q-8 .. 4q-1; the next pair is on its way
        kBase = k0;
        auto fetch_word = [&]() -> u32 {              // word kBase lives in slot kBase & 7
            u32 const v = lds_u32(ringLane + (kBase & 7) * (THREADS * 4));
            kBase++;
            if (!(kBase & 7)) store_next();           // wrapped to slot 0 = everything fetched: the next chunk goes there
            return v;
        };
        w0 = fetch_word(); w1 = fetch_word(); w2 = fetch_word();
        top_up(4);                                    // bring the ring to >= 5 unread words
    }

    // ---- decode ----
    u32 const tblCol = (u32)__cvta_generic_to_shared(tbl) + col * 2;       // + row * 128
    u32 const M = fx.mbits[col];
    u32 const shX = 32 - (tl ? tl : 1);                // window >> shX = index at full resolution
    u32 const shY = 32 - (M ? M : 1);                // umulhi(window, mulY) = window >> (32 - M) = first-level index
    int const dOff = (int)fx.dOff[col];
    u32 const tblColD = tblCol + (u32)dOff * (G * 2);  // row = min(x - D, y) + D: the "+ D" lives in the base address
    bool const hardBlk = fx.hard[col] != 0;

    static_assert(THREADS * 4 == 1024, "slot stride: (r >> 5) * 1024 = (r * 32) & ~1023");
    u32 const tid4 = tid * 4;
    u32 const ringBase = (u32)__cvta_generic_to_shared(ringRaw);
    u32 const slotBias = kBase * (THREADS * 4);        // (kBase + (r >> 5)) * 1024, masked to the 8 slots: the low 10 bits of r * 32 fall to the mask
    u32 hi = __funnelshift_l(w1, w0, r), lo = __funnelshift_l(w2, w1, r);
    // Row of the window: min(x, D + y) = min(x - D, y) + D, as a signed minimum.  Plain shifts: a high multiply for one of them
    // (to move work from the busy ALU pipe to the FMA pipe) measured 0.98 instead of 0.94 ms per GiB, both as IMAD.HI throttled
    // that pipe outright (round 2 v1).
#define HUFD_LOOKUP(E, H) do { \
        int const a_ = (int)((H) >> shX) - dOff; \
        int const y_ = (int)((H) >> shY); \
        E = lds_u16((u32)min(a_, y_) * (G * 2) + tblColD); \
    } while (0)
    // After a pair of symbols: `r` has grown by their bits (it is never reduced: the funnel shifts take it modulo 32 and
    // bit 5 FLIPS exactly when a 32-bit word has been used up -- a pair is at most 24 bits).  Then the words rotate and the
    // next one comes out of the ring, all predicated; the window is recomputed from the raw words either way.  The slot of that
    // next word follows from `r` itself -- word kBase + (r >> 5), slot = index mod 8 -- as one multiply-add (FMA pipe) and one
    // AND-OR with the lane's column offset; the ring sits at the start of shared memory, so its base is an immediate of the load.
    // `r` must stay a clean bit count for that: the pair's two lengths are added by one byte-wise dot product of the packed
    // cells (len0 | sym0 << 8 | len1 << 16 | sym1 << 24) with 0x00010001 -- also on the FMA pipe.
#define HUFD_ADVANCE(PACKED) do { \
        u32 const rOld_ = r; \
        r = __dp4a((u32)(PACKED), 0x00010001u, r); \
        u32 const ra_ = ((rOld_ * 32u * (THREADS * 4 / 1024u) + slotBias) & ((RW - 1) * THREADS * 4)) | tid4; \
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t" \
                     "xor.b32 t, %4, %5;\n\tand.b32 t, t, 32;\n\tsetp.ne.u32 p, t, 0;\n\t" \
                     "@p mov.b32 %0, %1;\n\t@p mov.b32 %1, %2;\n\t@p ld.shared.u32 %2, [%3];\n\t}" \
                     : "+r"(w0), "+r"(w1), "+r"(w2) : "r"(ringBase + ra_), "r"(r), "r"(rOld_) : "memory"); \
        hi = __funnelshift_l(w1, w0, r); lo = __funnelshift_l(w2, w1, r); \
    } while (0)

    u32 pos = 0;
    // 32 symbols -> ONE 32-byte store per lane (STG.256).  Every lane writes its own stream, so stores cannot coalesce across
    // lanes; what matters is that each one covers a whole 32-byte sector: measured on this access pattern (scripts/ubench/
    // scatter.cu) 16-byte stores cost 0.78 ms per GiB on their own -- partial-sector writes -- against 0.23 ms for 32-byte ones.
    bool const fastOk = go && !hardBlk && ((reinterpret_cast<u64>(outp) & 31) == 0);
    if (fastOk) {
        u32 const nIter = segLen >> 5;
        for (u32 it = 0; it < nIter; it++) {
            u32 o[8];
            #pragma unroll
            for (int h = 0; h < 8; h++) {
                if ((h & 3) == 0) top_up(4);                                     // group start: the common store
                else if ((h & 3) == 2 && __builtin_expect(__any_sync(__activemask(), unread() <= 3), 0)) top_up(3);   // mid-group: only when a lane ran low (lanes leave this loop at different trip counts: vote among those still in it)
                u32 e0, e1, e2, e3, hi1;
                HUFD_LOOKUP(e0, hi); hi1 = __funnelshift_l(lo, hi, e0); HUFD_LOOKUP(e1, hi1);
                u32 const p01 = e0 | (e1 << 16);
                HUFD_ADVANCE(p01);
                HUFD_LOOKUP(e2, hi); hi1 = __funnelshift_l(lo, hi, e2); HUFD_LOOKUP(e3, hi1);
                u32 const p23 = e2 | (e3 << 16);
                HUFD_ADVANCE(p23);
                o[h] = __byte_perm(p01, p23, 0x7531);
            }
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         :: "l"(outp + pos), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
            pos += 32;
        }
    }
    // ragged tails, unaligned segments and hard blocks: one symbol at a time
    if (pos < segLen) {
        const u16* const colp = tbl + col;
        auto parked_cell = [&](u32 idx) -> u32 {
            u32 w = 1;
            while (idx >= colp[(PARK_RANKEND + w) * G]) w++;
            u32 const first = (w == 1) ? 0u : colp[(PARK_RANKEND + w - 1) * G];
            u32 const k = colp[(PARK_LISTSTART + w) * G] + ((idx - first) >> (w - 1));
            return (tl + 1 - w) | (((colp[(k >> 1) * G] >> (8 * (k & 1))) & 0xFFu) << 8);
        };
        u32 cnt = 0;
        while (pos < segLen) {
            if ((cnt++ & 7) == 0) top_up(4);
            u32 e;
            if (hardBlk) e = parked_cell(hi >> (32 - tl));
            else HUFD_LOOKUP(e, hi);
            HUFD_ADVANCE(e);
            outp[pos++] = (u8)(e >> 8);
        }
    }
#undef HUFD_LOOKUP
#undef HUFD_ADVANCE

    // ---- verdict: every stream must be consumed exactly (huf_decompress.c:348-349) ----
    if (go) {
        u64 const fetched = 4ull * q - unread();                        // absolute index of the next word to fetch; w0 is word fetched-3
        u64 const consumed = 32ull * (fetched - 3) + (r & 31);
        if (consumed != (u64)expectBits) atomicMin(&fx.status[col], (u32)(5u << 8 | E_CORRUPT));
    }
    __syncthreads();
    if (tid < G) {
        u32 const bb = fx.bid[tid];
        if (bb != NOBLOCK && fx.kind[tid] == 0) {
            u32 const st = fx.status[tid];
            u64 rv = (st == NOERR) ? (u64)block_len(g, bb) : err(st & 0xFF);
            // Rejected only by the exact-consumption rule of the single-symbol decoder: where the reference would have run
            // its double-symbol decoder (always for HUF_decompress4X2, by HUF_selectDecoder for HUF_decompress) the verdict
            // is the second pass's (huf_x2_fixup.cu).
            if ((st >> 8) == 5u && ((flags & 2u) || (!(flags & 1u) && d_huf_select_decoder(block_len(g, bb), csizes[bb])))) rv = HUF_X2_PENDING;
            results[bb] = rv;
        }
    }
    // ---- raw / RLE blocks (huf_decompress.c:1065-1066; the harness' own 0-size convention, bench.c:393-402) ----
    for (int j = 0; j < G; j++) {
        int const kd = fx.kind[j];
        if (kd != 1 && kd != 2) continue;
        u32 const bb = fx.bid[j];
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

cudaError_t launch_huf_x2_fixup(const BatchGeom& g, void* dst, const void* cbuf, const u64* csizes, u64* results, cudaStream_t stream);

// flags: bit 0 = every block is a Huffman block (HUF_decompress4X1 / 4X2 semantics: no raw / RLE forms);
//        bit 1 = verdicts of the double-symbol decoder for every block (HUF_decompress4X2);
//        0     = HUF_decompress: raw / RLE forms, decoder (and hence verdict on malformed input) by HUF_selectDecoder.
cudaError_t launch_huf_decode(const BatchGeom& g, void* dst, const void* cbuf, const u64* csizes, u64* results,
                              const void* orig, cudaStream_t stream, u32 flags)
{
    static SmemOptIn optin;
    // Table rows per CTA.  Pass A: 370 rows -> 55.75 KB -> FOUR CTAs (256 blocks) per SM; pass B, for the blocks whose tables
    // need more (wide, flat alphabets: probagen P <= 10%): 808 rows -> two CTAs per SM.  FSEB200_HUFD_ROWS / _ROWS_B override.
    static u32 const rowsA = [] { const char* e = std::getenv("FSEB200_HUFD_ROWS"); u32 v = e ? (u32)std::atoi(e) : 370u; return v < hufd::MIN_ROWS ? hufd::MIN_ROWS : v > 1700u ? 1700u : v; }();
    static u32 const rowsB = [] { const char* e = std::getenv("FSEB200_HUFD_ROWS_B"); u32 v = e ? (u32)std::atoi(e) : 808u; return v > 1700u ? 1700u : v; }();   // 0 = single pass
    if (g.nBlocks == 0) return cudaSuccess;
    int const dev = current_device();
    size_t const smemA = hufd::smem_bytes(rowsA);
    cudaError_t e = optin.ensure(hufd::huf_decode_kernel, dev, (int)(rowsB > rowsA ? hufd::smem_bytes(rowsB) : smemA));
    if (e != cudaSuccess) return e;
    bool const twoPass = rowsB > rowsA;
    u32* scratch = nullptr;
    if (twoPass) {
        scratch = (u32*)stream_scratch(1, stream, ((size_t)g.nBlocks + 1) * sizeof(u32), &e);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(scratch, 0, sizeof(u32), stream);              // [0] = number of deferred blocks, [1..] = their indices
        if (e != cudaSuccess) return e;
    }
    // Pass A grid shape.  Every lane decodes a whole stream, so a launch lasts one "round" however few blocks a CTA holds, and a
    // round is the faster the fewer warps share an SM (2 warps per scheduler: ~0.35 ms; 7-8: ~0.95 ms).  A batch that fills most of
    // the machine is spread evenly -- every SM the same number of blocks per round, CTAs only partly full; a small batch (a pipeline
    // chunk, a scatter/gather piece, one block) packs its CTAs full instead, so that each SM hosts as few warps as possible.
    int perSm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, hufd::huf_decode_kernel, hufd::THREADS, smemA) != cudaSuccess || perSm < 1) perSm = 1;
    u32 const slots = (u32)perSm * (u32)device_sm_count(dev);
    u32 gEff = (u32)hufd::G;
    if ((u64)g.nBlocks * 5 > (u64)slots * hufd::G * 4) {                 // more than 80 % of what the machine holds at once (measured: 512 MiB packed 0.59 ms, 640 MiB spread 0.89 ms, 1 GiB spread 0.96 ms)
        u32 const rounds = (g.nBlocks + slots * hufd::G - 1) / (slots * hufd::G);
        gEff = (g.nBlocks + slots * rounds - 1) / (slots * rounds);
        if (gEff > (u32)hufd::G) gEff = hufd::G;
        if (gEff < 1) gEff = 1;
    }
    unsigned const grid = (g.nBlocks + gEff - 1) / gEff;
    hufd::huf_decode_kernel<<<grid, hufd::THREADS, smemA, stream>>>(g, (u8*)dst, (const u8*)cbuf, csizes, results, (const u8*)orig, flags, gEff, rowsA,
                                                                    nullptr, nullptr, twoPass ? scratch + 1 : nullptr, scratch);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (twoPass) {   // pass B over the deferred list; its length lives on the device, so the grid covers the worst case and idle CTAs leave at once
        unsigned const gridB = (g.nBlocks + hufd::G - 1) / hufd::G;
        hufd::huf_decode_kernel<<<gridB, hufd::THREADS, hufd::smem_bytes(rowsB), stream>>>(g, (u8*)dst, (const u8*)cbuf, csizes, results, (const u8*)orig, flags,
                                                                                            (u32)hufd::G, rowsB, scratch + 1, scratch, nullptr, nullptr);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (flags == 1u) return cudaSuccess;                                // X1-only semantics: no verdict pass
    return launch_huf_x2_fixup(g, dst, cbuf, csizes, results, stream);
}

}  // namespace fseb
