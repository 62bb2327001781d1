// huf_encode.cu -- batched Huff0 4-stream encode for sm_100a.
//
// Replaces, per block, the CPU chain
//   HUF_compress2 / HUF_compress_internal  lib/huf_compress.c:637-724,787-793 (no table reuse)
//   HIST_count_wksp                        lib/hist.c:163-173
//   HUF_optimalTableLog                    lib/huf_compress.c:48-51
//   HUF_buildCTable_wksp (+HUF_sort, HUF_setMaxHeight)  lib/huf_compress.c:215-410
//   HUF_writeCTable (+HUF_compressWeights) lib/huf_compress.c:63-147
//   HUF_compress4X_usingCTable_internal / HUF_compress1X_usingCTable_internal_body  :457-502,552-603
// and returns the same value per block: 0 (not compressible / does not fit), 1 (RLE, byte in dst[0]),
// the compressed size, or an error code.  Compressed bytes are identical to the reference's.
//
// B200 mapping -- two kernels, because the work has two very different shapes:
//   1. huf_plan_kernel, ONE WARP PER BLOCK: histogram (per-warp shared-memory counters), ranks by
//      counting, then the O(alphabet) serial pieces with exact CPU tie-breaks (Huffman merge, depth
//      limiter, canonical values, weight header incl. its tiny FSE coder) on lane 0.  Those chains are
//      latency-bound; with 24 blocks resident per SM they overlap each other instead of stalling a
//      whole CTA behind one lane (the first version did that: 88% of its stalls were barrier waits).
//      Output: a 1.25 KB "plan" per block in a stream-ordered scratch buffer (code table, header,
//      verdict).
//   2. huf_emit_kernel, ONE CTA (4 warps) PER BLOCK, one warp per stream, no serial section and no
//      second look at the data: stream sizes and offsets are already known from the per-segment
//      histograms of the plan kernel.  (An earlier version gave each thread a contiguous run of
//      symbols staged in shared memory: runs are 128 bytes apart, i.e. all 32 lanes of a warp in the
//      same bank -- 525M bank conflicts per 256 MiB, short-scoreboard 38 stalls per issue.)
#include <cstdlib>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "launch_util.cuh"
#include "fse_dev.cuh"
#include "sink_dev.cuh"
#include "huf_build_dev.cuh"

namespace fseb {
namespace hufe {

constexpr unsigned FULL = 0xFFFFFFFFu;

struct __align__(16) Plan {        // per block, in global scratch
    u32 ctable[256];               // val | nbBits << 16
    u8  header[136];
    u32 hSize;
    u32 state;                     // 0 = emit, 1 = verdict is final (nothing to emit)
    u32 total;                     // compressed size when state == 0
    u32 streamOff[4];              // byte offset of each stream from the start of the block
    u32 streamBytes[4];
};

// ---------------------------------------------------------------------------------------------
// kernel 1: statistics, code table, tree header, stream sizes, verdict -- one warp per block
// ---------------------------------------------------------------------------------------------
constexpr int PLAN_WARPS = 8;
struct PlanWarp {                  // 10.5 KB per warp, eight warps per CTA, two CTAs per SM.  More blocks in flight do not help: four warps per
                                   // CTA (five CTAs, 20 blocks per SM) measured 1.41 instead of 1.37 ms per GiB for the pair -- the kernel is
                                   // bound by its shared-memory atomics (LSU 55 %), not by the latency of the serial chains.  Packing two
                                   // 16-bit counters per word (24 in flight) loses the warp-aggregated ATOMS.POPC.INC the compiler emits
                                   // for a constant increment of 1: 1.81 ms, and 3.76 ms on P80 (26 lanes on one counter).
    u32   count4[4][256];          // one histogram per stream: stream sizes follow from them without touching the data again
    u32   count[256];
    HNode nodes[2 * 256 + 2];      // tree nodes during HUF_buildCTable; the header writer's scratch (1.5 KB) afterwards
    u32   ctable[256];
    u32   firstVal[32];         // first code value per length + running per-length counters
    u8    lenOf[256];
    alignas(4) u8 header[136];
};
static_assert(sizeof(HNode) * (2 * 256 + 2) >= 384 * sizeof(u32), "header scratch must fit in the node array");
__device__ __forceinline__ u32 seg_count(const PlanWarp& w, u32 k, u32 i) { return w.count4[k][i]; }

// histogram of src[begin, end) into cnt[256] (shared memory), one warp
__device__ __forceinline__ void warp_hist_range(u32* cnt, const u8* s, u32 begin, u32 end, unsigned lane)
{
    u32 i = begin;
    u32 const head = min(end, (u32)((begin + 15) & ~15u));         // bytes up to 16-byte alignment of the OFFSET ...
    bool const vec = ((reinterpret_cast<u64>(s) & 15) == 0);       // ... which is alignment of the address when the block is aligned
    if (vec) {
        for (u32 k = i + lane; k < head; k += 32) atomicAdd(&cnt[s[k]], 1u);
        i = head;
        u32 const nvec = (end - i) / 16;
        const uint4* const gv = reinterpret_cast<const uint4*>(s + i);
        for (u32 v0 = 0; v0 < nvec; v0 += 128) {                   // four 16-byte loads in flight per lane before the first use
            uint4 x[4];
            #pragma unroll
            for (int h = 0; h < 4; h++) { u32 const vi = v0 + 32 * h + lane; x[h] = (vi < nvec) ? __ldg(gv + vi) : make_uint4(0, 0, 0, 0); }
            #pragma unroll
            for (int h = 0; h < 4; h++) {
                if (v0 + 32 * h + lane >= nvec) continue;
                u32 const wd[4] = { x[h].x, x[h].y, x[h].z, x[h].w };
                #pragma unroll
                for (int k = 0; k < 4; k++) {
                    u32 const y = wd[k];                             // one extract + one address + one shared atomic per byte
                    atomicAdd(&cnt[y & 0xFF], 1u); atomicAdd(&cnt[__byte_perm(y, 0, 0x4441)], 1u);
                    atomicAdd(&cnt[__byte_perm(y, 0, 0x4442)], 1u); atomicAdd(&cnt[y >> 24], 1u);
                }
            }
        }
        i += nvec * 16;
    }
    for (u32 k = i + lane; k < end; k += 32) atomicAdd(&cnt[s[k]], 1u);
}

// All four per-segment histograms of an aligned block whose segments are whole 2 KB batches (the 32 KB case): one
// software pipeline over the block, the next batch's four 16-byte loads are in flight while the current one is counted,
// on two alternating register sets (nothing touches a register with a load in flight).
__device__ __forceinline__ void warp_hist4_pipelined(u32 (*count4)[256], const u8* s, u32 n, u32 seg, unsigned lane)
{
    const uint4* const gv = reinterpret_cast<const uint4*>(s) + lane;
    u32 const B = n / 2048, perSeg = seg / 2048;
    auto load = [&](uint4 (&x)[4], u32 b) {
        #pragma unroll
        for (int h = 0; h < 4; h++) x[h] = __ldg(gv + b * 128 + 32 * h);
    };
    auto count = [&](const uint4 (&x)[4], u32 b) {
        u32* const cnt = count4[b / perSeg];
        #pragma unroll
        for (int h = 0; h < 4; h++) {
            u32 const wd[4] = { x[h].x, x[h].y, x[h].z, x[h].w };
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                u32 const y = wd[k];                                 // one extract + one address + one shared atomic per byte
                atomicAdd(&cnt[y & 0xFF], 1u); atomicAdd(&cnt[__byte_perm(y, 0, 0x4441)], 1u);
                atomicAdd(&cnt[__byte_perm(y, 0, 0x4442)], 1u); atomicAdd(&cnt[y >> 24], 1u);
            }
        }
    };
    uint4 xa[4], xb[4];
    load(xa, 0);
    #pragma unroll 1
    for (u32 b = 0; b < B; b += 2) {
        if (b + 1 < B) load(xb, b + 1);
        count(xa, b);
        if (b + 2 < B) load(xa, b + 2);
        if (b + 1 < B) count(xb, b + 1);
    }
}

__global__ void __launch_bounds__(32 * PLAN_WARPS)
huf_plan_kernel(BatchGeom g, u8* __restrict__ cbuf, u64* __restrict__ csizes, const u8* __restrict__ src,
                unsigned msvReq, unsigned tlogReq, Plan* __restrict__ plans, int serialHeader)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PlanWarp& w = reinterpret_cast<PlanWarp*>(smem_raw)[threadIdx.x >> 5];
    unsigned const lane = threadIdx.x & 31u;
    u32 const b = blockIdx.x * PLAN_WARPS + (threadIdx.x >> 5);
    if (b >= g.nBlocks) return;
    u32 const n = block_len(g, b);
    const u8* const s = src + (u64)b * g.blockSize;
    u8* const d = cbuf + (u64)b * g.slot;
    u64 const cap = g.slot;
    Plan& P = plans[b];
#define FSEB_FINAL(v) do { if (lane == 0) { P.state = 1; csizes[b] = (v); } return; } while (0)

    // argument checks of HUF_compress_internal (huf_compress.c:656-664), in its order
    if (!n) FSEB_FINAL(0);
    if (!cap) FSEB_FINAL(0);
    if (n > HUF_BLOCK_MAX) FSEB_FINAL(err(E_SRC_WRONG));
    if (tlogReq > HUF_MAX_TLOG) FSEB_FINAL(err(E_TLOG_TOO_LARGE));
    if (msvReq > HUF_MAX_SV) FSEB_FINAL(err(E_MSV_TOO_LARGE));
    unsigned const msvDecl = msvReq ? msvReq : HUF_MAX_SV;
    unsigned huffLog = tlogReq ? tlogReq : HUF_DEF_TLOG;

    // ---- histograms, one per 4X segment (HIST_count_wksp semantics for their sum, hist.c:163-173,128) ----
    u32 const seg = (n + 3) / 4;
    for (u32 i = lane; i < 4 * 256; i += 32) (&w.count4[0][0])[i] = 0;
    __syncwarp();
    if ((reinterpret_cast<u64>(s) & 15) == 0 && seg % 2048 == 0 && n == 4 * seg) warp_hist4_pipelined(w.count4, s, n, seg, lane);
    else {
        #pragma unroll 1
        for (u32 k = 0; k < 4; k++) {
            u32 const beg = min(k * seg, n), end = (k < 3) ? min((k + 1) * seg, n) : n;
            warp_hist_range(w.count4[k], s, beg, end, lane);
        }
    }
    __syncwarp();
    u32 top = 0, best = 0;
    for (u32 i = lane; i < 256; i += 32) {
        u32 const c = seg_count(w, 0, i) + seg_count(w, 1, i) + seg_count(w, 2, i) + seg_count(w, 3, i);
        w.count[i] = c;
        if (c) top = i; best = c > best ? c : best;
    }
    #pragma unroll
    for (int dlt = 16; dlt; dlt >>= 1) { top = max(top, __shfl_xor_sync(FULL, top, dlt)); best = max(best, __shfl_xor_sync(FULL, best, dlt)); }
    __syncwarp();
    if (msvDecl < 255 && top > msvDecl) FSEB_FINAL(err(E_MSV_TOO_SMALL));                         // hist.c:128
    if (best == n) { if (lane == 0) d[0] = s[0]; FSEB_FINAL(1); }                                // huf_compress.c:673
    if (best <= (n >> 7) + 4) FSEB_FINAL(0);                                                      // :674
    u32 const msv = top;

    // ---- Huffman table (huf_compress.c:691-699) ----
    huffLog = d_optimal_tablelog(huffLog, n, msv, 1);
    {   u64 const mb = warp_huf_build_ctable(w.ctable, w.count, msv, huffLog, w.nodes, w.lenOf, w.firstVal);
        if (is_err(mb)) FSEB_FINAL(mb);
        huffLog = (u32)mb;
    }
    // ---- tree header (huf_compress.c:703-716) ----
    __syncwarp();
    u32* const wksp = reinterpret_cast<u32*>(w.nodes);                  // the tree is done with its nodes
    u64 hs = 0;
    if (serialHeader) { if (lane == 0) hs = d_huf_write_ctable(w.header, cap < sizeof(w.header) ? cap : sizeof(w.header), w.ctable, msv, huffLog, wksp); hs = __shfl_sync(FULL, hs, 0); }
    else hs = warp_huf_write_ctable(w.header, cap < sizeof(w.header) ? cap : sizeof(w.header), w.ctable, msv, huffLog, wksp);
    if (is_err(hs)) FSEB_FINAL(hs);
    if (hs + 12 >= n) FSEB_FINAL(0);
    u64 const capLeft = cap - hs;
    if (capLeft < 6 + 1 + 1 + 1 + 8 || n < 12) FSEB_FINAL(0);                                     // :564-565
    // ---- stream sizes from the per-segment histograms; the writer's capacity rule stream after stream
    //      (:566-600, bitstream.h:190,246,258) and the final compressibility test (:625) ----
    u32 bits[4];
    #pragma unroll
    for (int k = 0; k < 4; k++) {
        u32 acc = 0;
        for (u32 i = lane; i <= msv; i += 32) acc += seg_count(w, k, i) * (w.ctable[i] >> 16);
        #pragma unroll
        for (int dlt = 16; dlt; dlt >>= 1) acc += __shfl_xor_sync(FULL, acc, dlt);
        bits[k] = acc;
    }
    u64 op = 6; bool fits = true;
    u32 offs[4], lens[4];
    #pragma unroll
    for (int t = 0; t < 4; t++) {
        u64 const capk = capLeft - op;
        u64 const tot = (u64)bits[t] + 1;                                    // + end mark
        if (fits && (capk <= 8 || (tot >> 3) >= capk - 8)) fits = false;
        offs[t] = (u32)(hs + op); lens[t] = (u32)((tot + 7) >> 3);
        op += lens[t];
    }
    u64 const total = hs + op;
    if (!fits || total >= (u64)n - 1) FSEB_FINAL(0);
    __syncwarp();
    for (u32 i = lane; i < 256; i += 32) P.ctable[i] = w.ctable[i];
    for (u32 i = lane; i < (u32)hs; i += 32) P.header[i] = w.header[i];
    if (lane < 4) { P.streamOff[lane] = offs[lane]; P.streamBytes[lane] = lens[lane]; }
    if (lane == 0) { P.hSize = (u32)hs; P.state = 0; P.total = (u32)total; csizes[b] = total; }
#undef FSEB_FINAL
}

// ---------------------------------------------------------------------------------------------
// kernel 2: emit -- one CTA (4 warps) per block, one warp per stream, no serial section.
// A stream is the concatenation, last symbol first, of the codes.  The warp walks its segment from the
// end, 512 symbols at a time on aligned data: lane l takes the 8 symbols of one 64-bit piece of each of two
// 256-symbol groups (coalesced global reads, nothing staged), concatenates their codes (<= 88 bits per group),
// ONE warp scan of the two bit counts (packed in one register) gives both bit offsets, and the lane ORs its bits
// into a circular shared-memory window of the stream (aligned 32-bit words of the destination, whose position is
// known from the plan) with four reds at consecutive words; completed words leave 128 at a time, 16 bytes per lane.
// What bounds it was measured, not assumed: shared-memory wavefronts first (8-byte code cells), the ALU pipe now.
// ---------------------------------------------------------------------------------------------
constexpr int THREADS = 128;

// A 16-byte piece of a stream window that touches the stream's first or last bytes: only local bytes [a0, endByte) belong to
// this stream (the neighbours are another warp's), so it goes out byte by byte.  Out of line: once or twice per stream.
__device__ __noinline__ void store_piece_exact(u32* gw, u32 j, uint4 v, u32 a0, u32 endByte)
{
    u32 const wd[4] = { v.x, v.y, v.z, v.w };
    u8* const pb = reinterpret_cast<u8*>(gw + j);
    for (u32 t = 0; t < 16; t++) { u32 const at = 4 * j + t; if (at >= a0 && at < endByte) pb[t] = (u8)(wd[t >> 2] >> (8 * (t & 3))); }
}

__global__ void __launch_bounds__(THREADS)
huf_emit_kernel(BatchGeom g, u8* __restrict__ cbuf, const u8* __restrict__ src, const Plan* __restrict__ plans, const u32* __restrict__ sharedCT)
{
    constexpr u32 W = 512;                                          // words of stream window per warp (a double group adds <= 176, < 128 wait for the next flush)
    __shared__ u32 ctab[256];                                       // cells nbBits | code << 8.  4-byte cells: the kernel is bound by shared-memory
                                                                    // wavefronts (ncu: 95 % of the data pipe), and an 8-byte cell per lane costs two
    constexpr u32 WP = W + 4;                                       // + 3 spill words (a put writes up to 4 consecutive words, never wrapping) + 1 unused
    __shared__ __align__(16) u32 winAll[4 * WP];
    int const tid = threadIdx.x;
    unsigned const lane = tid & 31u; int const k = tid >> 5;
    u32 const b = blockIdx.x;
    const Plan& P = plans[b];
    if (P.state != 0) return;                                       // verdict already delivered by the plan kernel
    u32 const n = block_len(g, b);
    const u8* const s = src + (u64)b * g.blockSize;
    u8* const d = cbuf + (u64)b * g.slot;
    u32 const hSize = P.hSize, total = P.total;

    {   const u32* const ct = sharedCT ? sharedCT : P.ctable;       // one caller-supplied table for the whole batch, or the block's own
        u32 const c0 = ct[tid], c1 = ct[tid + 128];
        ctab[tid] = (c0 >> 16) | (c0 << 8 & 0xFFFF00u); ctab[tid + 128] = (c1 >> 16) | (c1 << 8 & 0xFFFF00u);
    }
    for (u32 i = tid; i < 4 * WP; i += THREADS) winAll[i] = 0;
    // tree header and jump table straight to the block (byte stores: the word they end in is shared with stream 1)
    for (u32 i = tid; i < hSize; i += THREADS) d[i] = P.header[i];
    if (tid < 3) { u32 const v = P.streamBytes[tid]; d[hSize + 2 * tid] = (u8)v; d[hSize + 2 * tid + 1] = (u8)(v >> 8); }
    __syncthreads();
    {
        u32 const seg = (n + 3) / 4;
        int const segBeg = (int)(k * seg);
        int const segEnd = (k < 3) ? (int)((k + 1) * seg) : (int)n;
        u32 sTab = (u32)__cvta_generic_to_shared(ctab);
        asm volatile("mov.u32 %0, %0;" : "+r"(sTab));               // keep it in a register (otherwise re-derived from the CTA id before every look-up)
        // The stream is built in a circular window of aligned 32-bit words of the destination and flushed as it grows, so a
        // CTA needs 6 KB of shared memory instead of an image of the whole block (occupancy: 10+ CTAs per SM instead of 6).
        u32* const win = winAll + k * WP;
        u32 sWin = (u32)__cvta_generic_to_shared(win);
        asm volatile("mov.u32 %0, %0;" : "+r"(sWin));
        u8* const gstart = d + P.streamOff[k];
        u32 const a0 = (u32)(reinterpret_cast<u64>(gstart) & 15);   // the window is 16-byte aligned in the destination: it is flushed in 16-byte pieces
        u32* const gw = reinterpret_cast<u32*>(gstart - a0);        // word j of the window <-> gw[j]
        u32 const sBytes = (k < 3) ? P.streamBytes[k] : total - P.streamOff[3];
        u32 const endByte = a0 + sBytes;                            // stream occupies local bytes [a0, endByte)
        u32 bitpos = 8u * a0;
        u32 flushed = 0;                                            // words already written out
        auto lds32 = [&](u32 a) -> u32 { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; };
        auto lds64 = [&](u32 o) -> uint2 { u32 const e = lds32(sTab + (o >> 1)); return make_uint2(e >> 8, e & 0xFFu); };   // {code, nbBits} of the cell at 8 * symbol
        auto red_or = [&](u32 a, u32 v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); };
        // three consecutive words, unconditionally: a predicated red compiles to a branch around it (4 more instructions each,
        // 162 instructions per 8 symbols in all), and a zero operand costs the shared-memory pipe no more than a one-lane red
        auto red_or3 = [&](u32 a, u32 v0, u32 v1, u32 v2) {
            asm volatile("red.shared.or.b32 [%0], %1; red.shared.or.b32 [%0+4], %2; red.shared.or.b32 [%0+8], %3;" :: "r"(a), "r"(v0), "r"(v1), "r"(v2) : "memory");
        };
        // warp scan of the lanes' bit counts: exclusive prefix in `excl`, returns the warp total
        auto scan = [&](u32 held, u32& excl) -> u32 {
            u32 incl = held;
            #pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1)                     // shfl.up's own predicate says whether a source lane exists
                asm volatile("{ .reg .pred p; .reg .u32 t; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.u32 %0, %0, t; }" : "+r"(incl) : "r"(dd));
            excl = incl - held;
            return __shfl_sync(FULL, incl, 31);
        };
        // ORs up to 64 bits (a0 | a1 << 32) into the window at bit offset `at`.  Only the first word's position wraps: words
        // W .. W+2 of the window are spill cells for words 0 .. 2 of the next lap, folded in when those are flushed.
        auto put = [&](u32 at, u32 x0, u32 x1) {
            u32 const sh = at & 31;
            red_or3(sWin + ((at >> 3) & (4 * (W - 1))), x0 << sh, __funnelshift_l(x0, x1, sh), __funnelshift_l(x1, 0, sh));
        };
        // the same for up to 88 bits (z0 | z1 << 32 | z2 << 64): four words
        auto put96 = [&](u32 at, u32 z0, u32 z1, u32 z2) {
            u32 const sh = at & 31;
            asm volatile("red.shared.or.b32 [%0], %1; red.shared.or.b32 [%0+4], %2; red.shared.or.b32 [%0+8], %3; red.shared.or.b32 [%0+12], %4;"
                         :: "r"(sWin + ((at >> 3) & (4 * (W - 1)))), "r"(z0 << sh), "r"(__funnelshift_l(z0, z1, sh)), "r"(__funnelshift_l(z1, z2, sh)),
                            "r"(__funnelshift_l(z2, 0, sh)) : "memory");
        };
        // Final words leave 128 at a time, one 16-byte piece per lane (the 32-words-per-flush version spent 50 of 135
        // instructions per 8 symbols here).  Between flushes < 128 final words wait, a double group adds <= 176 (+ 3 spill): < W.
        auto flush = [&](bool all) {
            u32 const upTo = all ? (endByte + 3) / 4 : (bitpos >> 5);
            if (!all && flushed + 128 > upTo) return;               // warp-uniform
            __syncwarp();
            while (all ? flushed < upTo : flushed + 128 <= upTo) {
                u32 const j = flushed + 4 * lane;
                if (j < upTo) {
                    u32 const i = j & (W - 1);
                    uint4 v = *reinterpret_cast<uint4*>(win + i);
                    *reinterpret_cast<uint4*>(win + i) = make_uint4(0, 0, 0, 0);
                    if (i == 0) { v.x |= win[W]; v.y |= win[W + 1]; v.z |= win[W + 2]; win[W] = 0; win[W + 1] = 0; win[W + 2] = 0; }
                    if (!all && (j != 0 || a0 == 0)) *reinterpret_cast<uint4*>(gw + j) = v;   // interior piece: below bitpos, inside the stream
                    else store_piece_exact(gw, j, v, a0, endByte);  // first / last pieces
                }
                flushed += 128;
            }
        };
        auto place = [&](u32 x0, u32 x1, u32 held) -> u32 {         // lane's `held` bits go to bitpos + exclusive prefix
            u32 excl; u32 const sum = scan(held, excl);
            put(bitpos + excl, x0, x1);
            return sum;
        };
        constexpr int PF = 4;
        int hiCur = segEnd;                                         // symbols [segBeg, hiCur) are still to be coded
        // ---- groups of 256 symbols on 8-byte-aligned data: lane l codes the 8 symbols of one 64-bit piece, highest byte first.
        //      One scan serves 8 symbols; the two 4-symbol halves (<= 44 bits each) are placed separately. ----
        {
            bool const al8 = ((reinterpret_cast<u64>(s) + (u64)segEnd) & 7) == 0;
            int const nBig = al8 ? (segEnd - segBeg) / 256 : 0;
            const uint2* gq = reinterpret_cast<const uint2*>(s + segEnd) - 1 - lane;         // group j: piece gq[-32 j]
            auto half = [&](u32 w, u32& a0, u32& a1) -> u32 {        // 4 symbols of one word -> up to 44 bits
                // a cell is used as a shift amount as it is: the funnel shift takes the low 5 bits (lengths <= 11, pairs <= 22)
                u32 const e0 = lds32(sTab + 4 * __byte_perm(w, 0, 0x4443)), e1 = lds32(sTab + 4 * __byte_perm(w, 0, 0x4442));
                u32 const e2 = lds32(sTab + 4 * __byte_perm(w, 0, 0x4441)), e3 = lds32(sTab + 4 * __byte_perm(w, 0, 0x4440));
                u32 const p01 = (e0 >> 8) | __funnelshift_l(0, e1 >> 8, e0), l01 = e0 + e1;
                u32 const p23 = (e2 >> 8) | __funnelshift_l(0, e3 >> 8, e2), l23 = e2 + e3;
                a0 = p01 | __funnelshift_l(0, p23, l01); a1 = __funnelshift_l(p23, 0, l01);
                return (l01 + l23) & 0xFFu;
            };
            // Z = X | Y << lx for two halves (X < 2^lx, lx <= 44, Y < 2^44): the lane's 8 symbols as one string of <= 88 bits, so
            // that they cost 4 window ORs instead of 6 -- the kernel is bound by shared-memory wavefronts, not by instructions
            auto merge = [&](u32 x0, u32 x1, u32 lx, u32 y0, u32 y1, u32& z0, u32& z1, u32& z2) {
                u32 const t0 = __funnelshift_l(0, y0, lx), t1 = __funnelshift_l(y0, y1, lx), t2 = __funnelshift_l(y1, 0, lx);   // Y << (lx & 31)
                bool const far = (lx & 32u) != 0;
                z0 = x0 | (far ? 0u : t0); z1 = x1 | (far ? t0 : t1); z2 = far ? t1 : t2;
            };
            // Two groups of 256 symbols per scan: the two bit counts ride in the halves of one register (each total < 2^16).
            auto big2 = [&](uint2 curA, uint2 curB) {
                u32 x0, x1, y0, y1, za0, za1, za2, zb0, zb1, zb2;
                u32 lx = half(curA.y, x0, x1), ly = half(curA.x, y0, y1);           // .y holds the higher addresses: emitted first
                merge(x0, x1, lx, y0, y1, za0, za1, za2);
                u32 const la = lx + ly;
                lx = half(curB.y, x0, x1); ly = half(curB.x, y0, y1);
                merge(x0, x1, lx, y0, y1, zb0, zb1, zb2);
                u32 const lb = lx + ly;
                u32 excl; u32 const sum = scan(la | (lb << 16), excl);
                u32 const sumA = sum & 0xFFFFu;
                put96(bitpos + (excl & 0xFFFFu), za0, za1, za2);
                put96(bitpos + sumA + (excl >> 16), zb0, zb1, zb2);
                bitpos += sumA + (sum >> 16);
                flush(false);
            };
            constexpr int PB = 2;                                   // pieces in flight per register set (2 x 256 B per warp)
            int const rounds = nBig / PB;
            uint2 bufA[PB], bufB[PB];
            #pragma unroll
            for (int i = 0; i < PB; i++) { bufA[i] = rounds > 0 ? __ldg(gq - 32 * i) : make_uint2(0, 0); bufB[i] = make_uint2(0, 0); }
            auto round = [&](uint2 (&X)[PB], uint2 (&Y)[PB], int r) {
                gq -= 32 * PB;
                uint2 const c0 = X[0], c1 = X[1];
                if (r + 1 < rounds) { Y[0] = __ldg(gq); Y[1] = __ldg(gq - 32); }
                big2(c0, c1);
            };
            int r = 0;
            #pragma unroll 1
            for (; r + 1 < rounds; r += 2) { round(bufA, bufB, r); round(bufB, bufA, r + 1); }
            if (r < rounds) round(bufA, bufB, r);
            hiCur -= 256 * rounds * PB;
        }
        // ---- full groups of 128 symbols on word-aligned data: lane l codes the 4 symbols of one 32-bit word, highest byte first ----
        bool const wordAligned = ((reinterpret_cast<u64>(s) + (u64)hiCur) & 3) == 0;
        int const nFull = wordAligned ? (hiCur - segBeg) / 128 : 0;
        {
            const u32* gp = reinterpret_cast<const u32*>(s + hiCur) - 1 - lane;             // group j: word gp[-32 j]
            auto group = [&](u32 o0, u32 o1, u32 o2, u32 o3) {      // table offsets of the 4 symbols, emission order
                uint2 const e0 = lds64(o0), e1 = lds64(o1), e2 = lds64(o2), e3 = lds64(o3);
                u32 const p01 = e0.x | (e1.x << e0.y), l01 = e0.y + e1.y;           // <= 22 bits
                u32 const p23 = e2.x | (e3.x << e2.y), l23 = e2.y + e3.y;
                u32 const a0 = p01 | (p23 << l01), a1 = __funnelshift_l(p23, 0, l01);
                bitpos += place(a0, a1, l01 + l23);
                flush(false);
            };
            // rounds of PF groups: a raw word is unpacked one round after its load was issued, and it is dead (unpacked into
            // table offsets) before the next load is issued into its register -- warps issue in order, so no instruction
            // may touch a register with a load in flight
            int const rounds = nFull / PF;
            u32 bufA[PF], bufB[PF];                                 // two register sets, used alternately: nothing is ever rotated
            #pragma unroll
            for (int i = 0; i < PF; i++) { bufA[i] = rounds > 0 ? __ldg(gp - 32 * i) : 0u; bufB[i] = 0u; }
            auto round = [&](u32 (&X)[PF], u32 (&Y)[PF], int r) {   // code the groups held in X, request the next round into Y
                gp -= 32 * PF;
                bool const more = r + 1 < rounds;
                #pragma unroll
                for (int i = 0; i < PF; i++) {
                    u32 const cur = X[i];
                    if (more) Y[i] = __ldg(gp - 32 * i);
                    group(8 * (cur >> 24), 8 * __byte_perm(cur, 0, 0x4442), 8 * __byte_perm(cur, 0, 0x4441), 8 * (cur & 0xFFu));
                }
            };
            int r = 0;
            #pragma unroll 1
            for (; r + 1 < rounds; r += 2) { round(bufA, bufB, r); round(bufB, bufA, r + 1); }
            if (r < rounds) round(bufA, bufB, r);
            for (int j = rounds * PF; j < nFull; j++) {             // < PF groups left
                u32 const cur = __ldg(reinterpret_cast<const u32*>(s + hiCur) - 1 - lane - 32 * j);
                group(8 * (cur >> 24), 8 * __byte_perm(cur, 0, 0x4442), 8 * __byte_perm(cur, 0, 0x4441), 8 * (cur & 0xFFu));
            }
        }
        // ---- what is left (a partial group, or everything when the segment end is not word aligned) ----
        {
            int const rest = hiCur - 128 * nFull;
            auto fetch = [&](int hi) -> u32 {                       // the 4 symbols below hi-4*lane as one little-endian word
                int const top = hi - 4 * (int)lane;                 // exclusive
                u32 v = 0;
                if (top - 4 >= segBeg && ((reinterpret_cast<u64>(s + top) & 3) == 0)) v = __ldg(reinterpret_cast<const u32*>(s + top - 4));
                else if (top > segBeg) {
                    #pragma unroll
                    for (int j = 0; j < 4; j++) { int const i = top - 1 - j; if (i >= segBeg) v |= (u32)s[i] << (8 * (3 - j)); }
                }
                return v;
            };
            #pragma unroll 1
            for (int hi = rest; hi > segBeg; hi -= 128) {
                u32 const cur = fetch(hi);
                int const nValid = hi - 4 * (int)lane - segBeg;     // symbols available to this lane
                u64 acc = 0; u32 held = 0;                          // up to 48 bits
                #pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint2 const e = lds64(8 * ((cur >> (8 * (3 - j))) & 0xFF));
                    if (j < nValid) { acc |= (u64)e.x << held; held += e.y; }
                }
                bitpos += place((u32)acc, (u32)(acc >> 32), held);
                flush(false);
            }
        }
        if (lane == 0) red_or(sWin + 4 * ((bitpos >> 5) & (W - 1)), 1u << (bitpos & 31));    // end mark (bitstream.h:256)
        bitpos += 1;
        flush(true);                                                // the rest, byte-exact at both ends
    }
}

// ---------------------------------------------------------------------------------------------
// Table reuse: every block of the batch coded with ONE caller-supplied HUF_CElt table, i.e. per block
// HUF_compress4X_usingCTable (lib/huf_compress.c:552-610).  No histogram, no tree, no header: the plan kernel disappears and
// this warp-per-block pass only sums the code lengths of each 4X segment (stream sizes, the writer's capacity rule).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * PLAN_WARPS)
huf_sizes_kernel(BatchGeom g, u64* __restrict__ csizes, const u8* __restrict__ src, const u32* __restrict__ ct, Plan* __restrict__ plans)
{
    __shared__ u8 nbOf[256];
    unsigned const lane = threadIdx.x & 31u;
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) nbOf[i] = (u8)(ct[i] >> 16);
    __syncthreads();
    u32 const b = blockIdx.x * PLAN_WARPS + (threadIdx.x >> 5);
    if (b >= g.nBlocks) return;
    u32 const n = block_len(g, b);
    const u8* const s = src + (u64)b * g.blockSize;
    u64 const cap = g.slot;
    Plan& P = plans[b];
    if (cap < 6 + 1 + 1 + 1 + 8 || n < 12) { if (lane == 0) { P.state = 1; csizes[b] = 0; } return; }   // huf_compress.c:564-565
    u32 const seg = (n + 3) / 4;
    u32 bits[4];
    #pragma unroll 1
    for (u32 k = 0; k < 4; k++) {
        u32 const beg = k * seg, end = (k < 3) ? (k + 1) * seg : n;
        u32 acc = 0;
        u32 i = beg + lane * 4;
        if ((reinterpret_cast<u64>(s + beg) & 3) == 0) {
            for (; i + 4 <= end; i += 128) {
                u32 const w = __ldg(reinterpret_cast<const u32*>(s + i));
                acc += nbOf[w & 0xFF] + nbOf[(w >> 8) & 0xFF] + nbOf[(w >> 16) & 0xFF] + nbOf[w >> 24];
            }
            for (u32 t = i; t < end && t < i + 4; t++) acc += nbOf[s[t]];      // the lane whose word straddles the end
        } else for (u32 t = beg + lane; t < end; t += 32) acc += nbOf[s[t]];
        #pragma unroll
        for (int dlt = 16; dlt; dlt >>= 1) acc += __shfl_xor_sync(FULL, acc, dlt);
        bits[k] = acc;
    }
    u64 op = 6; bool fits = true;
    u32 offs[4], lens[4];
    #pragma unroll
    for (int t = 0; t < 4; t++) {                                            // :566-600 with bitstream.h:190,246,258
        u64 const capk = cap - op;
        u64 const tot = (u64)bits[t] + 1;
        if (fits && (capk <= 8 || (tot >> 3) >= capk - 8)) fits = false;
        offs[t] = (u32)op; lens[t] = (u32)((tot + 7) >> 3);
        op += lens[t];
    }
    if (!fits) { if (lane == 0) { P.state = 1; csizes[b] = 0; } return; }
    if (lane < 4) { P.streamOff[lane] = offs[lane]; P.streamBytes[lane] = lens[lane]; }
    if (lane == 0) { P.hSize = 0; P.state = 0; P.total = (u32)op; csizes[b] = op; }
}

}  // namespace hufe

cudaError_t launch_huf_encode_using_ctable(const BatchGeom& g, void* cbuf, u64* csizes, const void* src, const u32* dCTable, cudaStream_t stream)
{
    if (g.nBlocks == 0) return cudaSuccess;
    cudaError_t e;
    hufe::Plan* const plans = (hufe::Plan*)stream_scratch(2, stream, sizeof(hufe::Plan) * (size_t)g.nBlocks, &e);
    if (e != cudaSuccess) return e;
    unsigned const grid = (g.nBlocks + hufe::PLAN_WARPS - 1) / hufe::PLAN_WARPS;
    hufe::huf_sizes_kernel<<<grid, 32 * hufe::PLAN_WARPS, 0, stream>>>(g, csizes, (const u8*)src, dCTable, plans);
    hufe::huf_emit_kernel<<<g.nBlocks, hufe::THREADS, 0, stream>>>(g, (u8*)cbuf, (const u8*)src, plans, dCTable);
    return cudaGetLastError();
}

cudaError_t launch_huf_encode(const BatchGeom& g, void* cbuf, u64* csizes, const void* src,
                              unsigned msv, unsigned tlog, cudaStream_t stream)
{
    if (g.nBlocks == 0) return cudaSuccess;
    cudaError_t e;
    hufe::Plan* plans = nullptr;
    // Plan scratch (1.25 KB per block).  Default: one grow-only buffer per stream, reused by successive calls (stream order
    // makes that safe; a different stream gets its own buffer).  FSEB200_SCRATCH_ASYNC=1 switches to cudaMallocAsync/FreeAsync.
    static int asyncScratch = -1;
    if (asyncScratch < 0) { const char* const v = getenv("FSEB200_SCRATCH_ASYNC"); asyncScratch = (v && atoi(v) == 1) ? 1 : 0; }
    size_t const need = sizeof(hufe::Plan) * (size_t)g.nBlocks;
    if (asyncScratch) e = cudaMallocAsync((void**)&plans, need, stream);
    else {
        struct Slot { int dev; cudaStream_t s; void* p; size_t cap; };
        static std::mutex mu; static std::vector<Slot> slots;
        std::lock_guard<std::mutex> lock(mu);
        int dev = 0; cudaGetDevice(&dev);                           // the legacy stream handle is the same on every device
        Slot* hit = nullptr;
        for (auto& sl : slots) if (sl.s == stream && sl.dev == dev) { hit = &sl; break; }
        if (!hit) { slots.push_back(Slot{ dev, stream, nullptr, 0 }); hit = &slots.back(); }
        e = cudaSuccess;
        if (hit->cap < need) {
            if (hit->p) { cudaStreamSynchronize(stream); cudaFree(hit->p); hit->p = nullptr; hit->cap = 0; }
            e = cudaMalloc(&hit->p, need);
            if (e == cudaSuccess) hit->cap = need;
        }
        plans = (hufe::Plan*)hit->p;
    }
    if (e != cudaSuccess) return e;
    static int serialHeader = -1;
    if (serialHeader < 0) { const char* const v = getenv("FSEB200_HUF_SERIAL_HEADER"); serialHeader = (v && atoi(v) == 1) ? 1 : 0; }   // tuning knob
    // The source is read twice, by the plan kernel (histogram) and by the emit kernel: 1.70x the algorithmic DRAM bytes.  Walking the
    // batch in sub-batches whose source fits the 126 MB L2 (FSEB200_HUF_ENC_SUBBATCH = blocks per sub-batch) turns the second read into
    // L2 hits, but every launch pair then pays the plan kernel's tail (its per-block serial chains drain with the SMs mostly idle):
    // measured on a B200, 1 GiB P14: one pair 1.37 ms; 2048-block sub-batches 1.77 ms; 1024-block 2.36 ms.  Both kernels are
    // issue-bound, not DRAM-bound, so the re-read costs nothing in time and the default (0) keeps one launch pair for the whole batch.
    static u32 const subBatch = [] { const char* const v = getenv("FSEB200_HUF_ENC_SUBBATCH"); long n = v ? atol(v) : 0; return (u32)(n < 0 ? 0 : n); }();
    size_t const smem = sizeof(hufe::PlanWarp) * hufe::PLAN_WARPS;
    static SmemOptIn optin;
    e = optin.ensure(hufe::huf_plan_kernel, current_device(), (int)smem);
    if (e != cudaSuccess) return e;
    u32 const step = (subBatch && subBatch < g.nBlocks) ? subBatch : g.nBlocks;
    for (u32 b0 = 0; b0 < g.nBlocks; b0 += step) {
        BatchGeom gs = g;
        gs.nBlocks = (g.nBlocks - b0 < step) ? g.nBlocks - b0 : step;
        u64 const off = (u64)b0 * g.blockSize;
        u64 const span = (u64)gs.nBlocks * g.blockSize;
        gs.total = (g.total - off < span) ? g.total - off : span;
        u8* const cb = (u8*)cbuf + (u64)b0 * g.slot; const u8* const sp = (const u8*)src + off;
        unsigned const grid = (gs.nBlocks + hufe::PLAN_WARPS - 1) / hufe::PLAN_WARPS;
        hufe::huf_plan_kernel<<<grid, 32 * hufe::PLAN_WARPS, smem, stream>>>(gs, cb, csizes + b0, sp, msv, tlog, plans + b0, serialHeader);
        hufe::huf_emit_kernel<<<gs.nBlocks, hufe::THREADS, 0, stream>>>(gs, cb, sp, plans + b0, nullptr);
    }
    e = cudaGetLastError();
    cudaError_t const e2 = asyncScratch ? cudaFreeAsync(plans, stream) : cudaSuccess;
    return e != cudaSuccess ? e : e2;
}

}  // namespace fseb
