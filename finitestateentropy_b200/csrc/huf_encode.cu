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
//   2. huf_emit_kernel, ONE CTA PER BLOCK, no serial section: the block is staged once into shared
//      memory; a Huff0 stream is the concatenation, last symbol first, of the codes, so every thread
//      owns a run of symbols, an exclusive scan of the run bit-lengths gives its bit offset, and it ORs
//      its codes into a shared-memory image of the whole compressed block, copied out with aligned
//      16-byte stores (image and destination share their alignment mod 16).
#include "common.cuh"
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
    u32 state;                     // 0 = emit, 1 = verdict is final
    u64 verdict;
};

// ---------------------------------------------------------------------------------------------
// kernel 1: statistics, code table, tree header -- one warp per block
// ---------------------------------------------------------------------------------------------
constexpr int PLAN_WARPS = 8;
struct PlanWarp {
    u32   count[256];
    HNode nodes[2 * 256 + 2];
    u32   ctable[256];
    u32   firstVal[16];
    u8    lenOf[256];
    u8    header[136];
    u32   wksp[384];
};

__global__ void __launch_bounds__(32 * PLAN_WARPS)
huf_plan_kernel(BatchGeom g, u8* __restrict__ cbuf, u64* __restrict__ csizes, const u8* __restrict__ src,
                unsigned msvReq, unsigned tlogReq, Plan* __restrict__ plans)
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
#define FSEB_FINAL(v) do { if (lane == 0) { P.state = 1; P.verdict = (v); csizes[b] = (v); } return; } while (0)

    // argument checks of HUF_compress_internal (huf_compress.c:656-664), in its order
    if (!n) FSEB_FINAL(0);
    if (!cap) FSEB_FINAL(0);
    if (n > HUF_BLOCK_MAX) FSEB_FINAL(err(E_SRC_WRONG));
    if (tlogReq > HUF_MAX_TLOG) FSEB_FINAL(err(E_TLOG_TOO_LARGE));
    if (msvReq > HUF_MAX_SV) FSEB_FINAL(err(E_MSV_TOO_LARGE));
    unsigned const msvDecl = msvReq ? msvReq : HUF_MAX_SV;
    unsigned huffLog = tlogReq ? tlogReq : HUF_DEF_TLOG;

    // ---- histogram (HIST_count_wksp semantics, hist.c:163-173,128) ----
    for (u32 i = lane; i < 256; i += 32) w.count[i] = 0;
    __syncwarp();
    {
        u32 done = 0;
        if ((reinterpret_cast<u64>(s) & 15) == 0) {
            u32 const nvec = n / 16;
            const uint4* const gv = reinterpret_cast<const uint4*>(s);
            for (u32 i = lane; i < nvec; i += 32) {
                uint4 const v = __ldg(gv + i);
                u32 const wd[4] = { v.x, v.y, v.z, v.w };
                #pragma unroll
                for (int k = 0; k < 4; k++) {
                    u32 const x = wd[k];
                    u32 const b0 = x & 0xFF, b1 = (x >> 8) & 0xFF, b2 = (x >> 16) & 0xFF, b3 = x >> 24;
                    if ((b0 == b1) & (b1 == b2) & (b2 == b3)) atomicAdd(&w.count[b0], 4u);
                    else { atomicAdd(&w.count[b0], 1u); atomicAdd(&w.count[b1], 1u); atomicAdd(&w.count[b2], 1u); atomicAdd(&w.count[b3], 1u); }
                }
            }
            done = nvec * 16;
        }
        for (u32 i = done + lane; i < n; i += 32) atomicAdd(&w.count[s[i]], 1u);
    }
    __syncwarp();
    u32 top = 0, best = 0;
    for (u32 i = lane; i < 256; i += 32) { u32 const c = w.count[i]; if (c) top = i; best = c > best ? c : best; }
    #pragma unroll
    for (int dlt = 16; dlt; dlt >>= 1) { top = max(top, __shfl_xor_sync(FULL, top, dlt)); best = max(best, __shfl_xor_sync(FULL, best, dlt)); }
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
    u64 hs = 0;
    if (lane == 0) hs = d_huf_write_ctable(w.header, cap < sizeof(w.header) ? cap : sizeof(w.header), w.ctable, msv, huffLog, w.wksp);
    hs = __shfl_sync(FULL, hs, 0);
    if (is_err(hs)) FSEB_FINAL(hs);
    if (hs + 12 >= n) FSEB_FINAL(0);
    if (cap - hs < 6 + 1 + 1 + 1 + 8 || n < 12) FSEB_FINAL(0);                                    // :564-565
    __syncwarp();
    for (u32 i = lane; i < 256; i += 32) P.ctable[i] = w.ctable[i];
    for (u32 i = lane; i < (u32)hs; i += 32) P.header[i] = w.header[i];
    if (lane == 0) { P.hSize = (u32)hs; P.state = 0; P.verdict = 0; }
#undef FSEB_FINAL
}

// ---------------------------------------------------------------------------------------------
// kernel 2: emit -- one CTA per block, everything parallel
// ---------------------------------------------------------------------------------------------
constexpr int THREADS = 256;

struct EmitShared {
    u32 ctable[256];
    u32 chunkBits[THREADS];
    u32 chunkOff[THREADS];
    u32 streamBytes[4];
    u32 streamOff[4];
    u32 flag, total;
};

template <bool STAGED>
__global__ void __launch_bounds__(THREADS)
huf_emit_kernel(BatchGeom g, u8* __restrict__ cbuf, u64* __restrict__ csizes, const u8* __restrict__ src,
                const Plan* __restrict__ plans, u32 stageBytes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EmitShared& sh = *reinterpret_cast<EmitShared*>(smem_raw);
    u8* const srcStage = smem_raw + ((sizeof(EmitShared) + 15) & ~(size_t)15);
    u32* const image = reinterpret_cast<u32*>(srcStage + stageBytes);
    int const tid = threadIdx.x;
    u32 const b = blockIdx.x;
    const Plan& P = plans[b];
    if (P.state != 0) return;                                       // verdict already delivered by the plan kernel
    u32 const n = block_len(g, b);
    const u8* const s = src + (u64)b * g.blockSize;
    u8* const d = cbuf + (u64)b * g.slot;
    u64 const cap = g.slot;
    u32 const hSize = P.hSize;
    u64 const capLeft = cap - hSize;

    sh.ctable[tid] = P.ctable[tid];
    if (tid == 0) sh.flag = 0;
    if (STAGED) {                                                   // HBM/L2 -> shared, once
        if ((reinterpret_cast<u64>(s) & 15) == 0) {
            u32 const nvec = n / 16;
            const uint4* const gv = reinterpret_cast<const uint4*>(s);
            uint4* const sv = reinterpret_cast<uint4*>(srcStage);
            for (u32 i = tid; i < nvec; i += THREADS) sv[i] = __ldg(gv + i);
            for (u32 i = nvec * 16 + tid; i < n; i += THREADS) srcStage[i] = s[i];
        } else {
            for (u32 i = tid; i < n; i += THREADS) srcStage[i] = s[i];
        }
    }
    __syncthreads();
    auto sym = [&](u32 i) -> u32 { return STAGED ? (u32)srcStage[i] : (u32)s[i]; };

    // ---- pass A: bit length of every thread's run; stream k is owned by threads 64k..64k+63 ----
    u32 const seg = (n + 3) / 4;
    int const k = tid >> 6, j = tid & 63;
    u32 const segBeg = (u32)k * seg;
    u32 const segEnd = (k < 3) ? segBeg + seg : n;
    u32 const segLen = segEnd - segBeg;
    u32 const run = (segLen + 63) / 64;
    u32 const offHi = min((u32)j * run, segLen), offLo = min((u32)(j + 1) * run, segLen);
    u32 const hiC = segEnd - offHi, loIdx = segEnd - offLo;        // this thread's run = symbols [loIdx, hiC), emitted high to low
    {   u32 bits = 0;
        for (u32 i = loIdx; i < hiC; i++) bits += sh.ctable[sym(i)] >> 16;
        sh.chunkBits[tid] = bits;
    }
    __syncthreads();
    {   // exclusive scan over the 64 runs of each stream (2 warps per stream)
        u32 const v = sh.chunkBits[tid];
        u32 incl = v;
        unsigned const lane = tid & 31;
        #pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { u32 const t = __shfl_up_sync(FULL, incl, dd); if (lane >= (unsigned)dd) incl += t; }
        if (lane == 31) sh.chunkOff[tid] = incl;                    // warp total, parked in its last slot
        __syncthreads();
        u32 const firstHalf = sh.chunkOff[64 * k + 31];
        u32 const base = (j >= 32) ? firstHalf : 0u;
        u32 const total = firstHalf + sh.chunkOff[64 * k + 63];
        __syncthreads();
        sh.chunkOff[tid] = base + incl - v;
        if (j == 0) sh.streamBytes[k] = total;                      // bits for now
    }
    __syncthreads();
    if (tid == 0) {                                // sizes with the writer's capacity rule, stream after stream (:566-600, bitstream.h:190,246,258)
        u64 op = 6; bool fits = true;
        for (int t = 0; t < 4 && fits; t++) {
            u64 const capk = capLeft - op;
            u64 const tot = (u64)sh.streamBytes[t] + 1;                       // + end mark
            if (capk <= 8 || (tot >> 3) >= capk - 8) { fits = false; break; }
            sh.streamOff[t] = (u32)op;
            sh.streamBytes[t] = (u32)((tot + 7) >> 3);
            op += sh.streamBytes[t];
        }
        u64 const total = hSize + op;
        if (!fits || total >= (u64)n - 1) { sh.flag = 1; csizes[b] = 0; }                            // :625
        else sh.total = (u32)total;
    }
    __syncthreads();
    if (sh.flag) return;
    u32 const total = sh.total;

    // ---- pass B: build the block image in shared memory ----
    u32 const al = (u32)(reinterpret_cast<u64>(d) & 15);          // image byte i <-> d[i - al]
    u32 const imgWords = (al + total + 3) / 4 + 1;
    for (u32 i = tid; i < imgWords; i += THREADS) image[i] = 0;
    __syncthreads();
    {   u8* const img8 = reinterpret_cast<u8*>(image);
        for (u32 i = tid; i < hSize; i += THREADS) img8[al + i] = P.header[i];
        if (tid < 3) { u32 const v = sh.streamBytes[tid]; img8[al + hSize + 2 * tid] = (u8)v; img8[al + hSize + 2 * tid + 1] = (u8)(v >> 8); }
    }
    __syncthreads();
    {   u64 const Pb = 8ull * (al + hSize + sh.streamOff[k]) + sh.chunkOff[tid];
        u32* wp = image + (Pb >> 5);
        unsigned held = (unsigned)(Pb & 31);
        u64 acc = 0;
        for (u32 i = hiC; i-- > loIdx;) {
            u32 const e = sh.ctable[sym(i)];
            acc |= (u64)(e & 0xFFFF) << held;
            held += e >> 16;
            if (held >= 32) { atomicOr(wp++, (u32)acc); acc >>= 32; held -= 32; }
        }
        if (loIdx == segBeg && hiC > loIdx) { acc |= 1ull << held; held++; if (held >= 32) { atomicOr(wp++, (u32)acc); acc >>= 32; held -= 32; } }   // end mark by the owner of the first symbol
        if (held) atomicOr(wp, (u32)acc);
    }
    __syncthreads();
    // ---- copy out (HBM write) ----
    {   const u8* const img8 = reinterpret_cast<const u8*>(image);
        u32 const first16 = (al + 15) & ~15u;
        u32 const endOff = al + total;
        if (first16 >= endOff) { for (u32 i = al + tid; i < endOff; i += THREADS) d[i - al] = img8[i]; }
        else {
            for (u32 i = al + tid; i < first16; i += THREADS) d[i - al] = img8[i];
            u32 const nvec = (endOff - first16) / 16;
            const uint4* const iv = reinterpret_cast<const uint4*>(img8 + first16);
            uint4* const ov = reinterpret_cast<uint4*>(d + (first16 - al));
            for (u32 i = tid; i < nvec; i += THREADS) ov[i] = iv[i];
            for (u32 i = first16 + nvec * 16 + tid; i < endOff; i += THREADS) d[i - al] = img8[i];
        }
    }
    if (tid == 0) csizes[b] = total;
}

}  // namespace hufe

cudaError_t launch_huf_encode(const BatchGeom& g, void* cbuf, u64* csizes, const void* src,
                              unsigned msv, unsigned tlog, cudaStream_t stream)
{
    if (g.nBlocks == 0) return cudaSuccess;
    cudaError_t e;
    hufe::Plan* plans = nullptr;
    e = cudaMallocAsync((void**)&plans, sizeof(hufe::Plan) * (size_t)g.nBlocks, stream);     // stream-ordered scratch
    if (e != cudaSuccess) return e;
    {   size_t const smem = sizeof(hufe::PlanWarp) * hufe::PLAN_WARPS;
        static bool configured = false;
        if (!configured) {
            e = cudaFuncSetAttribute(hufe::huf_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured = true;
        }
        unsigned const grid = (g.nBlocks + hufe::PLAN_WARPS - 1) / hufe::PLAN_WARPS;
        hufe::huf_plan_kernel<<<grid, 32 * hufe::PLAN_WARPS, smem, stream>>>(g, (u8*)cbuf, csizes, (const u8*)src, msv, tlog, plans);
    }
    {   bool const staged = g.blockSize <= 65536;                                              // 2 x 128 KB would not fit in 227 KB
        u32 const stageBytes = staged ? ((g.blockSize + 15u + 16u) & ~15u) : 0u;
        size_t const imageBytes = (size_t)g.blockSize + 64 + 16;                               // accepted blocks are < n bytes
        size_t const smem = ((sizeof(hufe::EmitShared) + 15) & ~(size_t)15) + stageBytes + imageBytes + 16;
        static size_t configured[2] = { 0, 0 };
        if (smem > configured[staged]) {
            e = staged ? cudaFuncSetAttribute(hufe::huf_emit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                       : cudaFuncSetAttribute(hufe::huf_emit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured[staged] = smem;
        }
        if (staged) hufe::huf_emit_kernel<true><<<g.nBlocks, hufe::THREADS, smem, stream>>>(g, (u8*)cbuf, csizes, (const u8*)src, plans, stageBytes);
        else hufe::huf_emit_kernel<false><<<g.nBlocks, hufe::THREADS, smem, stream>>>(g, (u8*)cbuf, csizes, (const u8*)src, plans, stageBytes);
    }
    e = cudaGetLastError();
    cudaError_t const e2 = cudaFreeAsync(plans, stream);
    return e != cudaSuccess ? e : e2;
}

}  // namespace fseb
