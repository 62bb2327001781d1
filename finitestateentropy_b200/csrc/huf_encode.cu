// huf_encode.cu -- batched Huff0 4-stream encode for sm_100a: one CTA per block.
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
// B200 mapping:
//   * the block is staged once into shared memory (coalesced 16-byte loads), so HBM sees it once:
//     the histogram and the encoder both read the staged copy;
//   * histogram: per-warp private shared-memory counters, merged by the CTA;
//   * table: ranks by counting (each thread ranks one symbol against the broadcast counts), the
//     O(alphabet) Huffman merge / depth limiter / header coder run on one lane (exact CPU tie-breaks);
//   * encode: a Huff0 stream is the concatenation, last symbol first, of the codes -> every thread
//     owns a run of symbols, an exclusive scan of the run bit-lengths gives its bit offset, and it
//     ORs its codes into a shared-memory image of the whole compressed block, which is then copied
//     out with aligned 16-byte stores (image and destination share their alignment mod 16).
#include "common.cuh"
#include "fse_dev.cuh"
#include "sink_dev.cuh"
#include "huf_build_dev.cuh"

namespace fseb {
namespace hufe {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;

struct Shared {                    // fixed part; the staged block and the output image follow in dynamic smem
    u32 count[256];
    u32 ctable[256];               // val | nbBits << 16
    HNode nodes[2 * 256 + 2];
    u32 chunkBits[THREADS];
    u32 chunkOff[THREADS];
    u32 streamBytes[4];
    u32 streamOff[4];
    u8  header[136];
    u64 verdict;                   // final return value once known
    u32 flag;                      // 0 = keep going, 1 = verdict final
    u32 hSize, maxBits, msv, largest, total;
    u32 wksp[384];                 // weight-header FSE scratch
};

__global__ void __launch_bounds__(THREADS)
huf_encode_kernel(BatchGeom g, u8* __restrict__ cbuf, u64* __restrict__ csizes, const u8* __restrict__ src,
                  unsigned msvReq, unsigned tlogReq, u32 stageBytes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
    u8* const srcStage = smem_raw + ((sizeof(Shared) + 15) & ~(size_t)15);
    u32* const image = reinterpret_cast<u32*>(srcStage + stageBytes);        // compressed-block image (also: warp histograms)
    int const tid = threadIdx.x, warp = tid >> 5;
    u32 const b = blockIdx.x;
    u32 const n = block_len(g, b);
    const u8* const s = src + (u64)b * g.blockSize;
    u8* const d = cbuf + (u64)b * g.slot;
    u64 const cap = g.slot;

    // ---- argument checks of HUF_compress_internal (huf_compress.c:656-664), in its order ----
    if (tid == 0) {
        sh.flag = 0; sh.verdict = 0;
        if (!n) { sh.flag = 1; sh.verdict = 0; }
        else if (!cap) { sh.flag = 1; sh.verdict = 0; }
        else if (n > HUF_BLOCK_MAX) { sh.flag = 1; sh.verdict = err(E_SRC_WRONG); }
        else if (tlogReq > HUF_MAX_TLOG) { sh.flag = 1; sh.verdict = err(E_TLOG_TOO_LARGE); }
        else if (msvReq > HUF_MAX_SV) { sh.flag = 1; sh.verdict = err(E_MSV_TOO_LARGE); }
    }
    __syncthreads();
    if (sh.flag) { if (tid == 0) csizes[b] = sh.verdict; return; }
    unsigned const msvDecl = msvReq ? msvReq : HUF_MAX_SV;
    unsigned huffLog = tlogReq ? tlogReq : HUF_DEF_TLOG;

    // ---- stage the block (HBM read #1 and only); blocks too large to stage are read in place ----
    bool const staged = stageBytes != 0;
    const u8* const sp = staged ? srcStage : s;
    if (staged) {
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
    // ---- histogram (HIST_count_wksp semantics, hist.c:163-173,128) ----
    u32* const whist = image;                                   // WARPS x 256 counters
    for (int i = tid; i < WARPS * 256; i += THREADS) whist[i] = 0;
    __syncthreads();
    {   u32* const mine = whist + warp * 256;
        const u32* const w32 = reinterpret_cast<const u32*>(srcStage);
        u32 const nw = staged ? n / 4 : 0;
        for (u32 i = tid; i < nw; i += THREADS) {
            u32 const v = w32[i];
            u32 const b0 = v & 0xFF, b1 = (v >> 8) & 0xFF, b2 = (v >> 16) & 0xFF, b3 = v >> 24;
            if ((b0 == b1) & (b1 == b2) & (b2 == b3)) atomicAdd(&mine[b0], 4u);
            else { atomicAdd(&mine[b0], 1u); atomicAdd(&mine[b1], 1u); atomicAdd(&mine[b2], 1u); atomicAdd(&mine[b3], 1u); }
        }
        for (u32 i = nw * 4 + tid; i < n; i += THREADS) atomicAdd(&mine[sp[i]], 1u);
    }
    __syncthreads();
    {   u32 c = 0;
        #pragma unroll
        for (int w = 0; w < WARPS; w++) c += whist[w * 256 + tid];
        sh.count[tid] = c;
    }
    __syncthreads();
    if (tid == 0) {
        u32 top = 255; while (!sh.count[top]) top--;
        u32 best = 0;
        for (u32 i = 0; i <= top; i++) best = sh.count[i] > best ? sh.count[i] : best;
        sh.msv = top; sh.largest = best;
        if (msvDecl < 255 && top > msvDecl) { sh.flag = 1; sh.verdict = err(E_MSV_TOO_SMALL); }     // hist.c:128
        else if (best == n) { d[0] = sp[0]; sh.flag = 1; sh.verdict = 1; }                     // huf_compress.c:673
        else if (best <= (n >> 7) + 4) { sh.flag = 1; sh.verdict = 0; }                              // :674
    }
    __syncthreads();
    if (sh.flag) { if (tid == 0) csizes[b] = sh.verdict; return; }
    u32 const msv = sh.msv;

    // ---- Huffman table (huf_compress.c:691-699) ----
    huffLog = d_optimal_tablelog(huffLog, n, msv, 1);
    {   u64 const mb = cta_huf_build_ctable(sh.ctable, sh.count, msv, huffLog, sh.nodes, sh.chunkBits, sh.chunkOff);
        if (is_err(mb)) { if (tid == 0) csizes[b] = mb; return; }
        huffLog = (u32)mb;
    }
    // ---- tree header (huf_compress.c:703-716) ----
    if (tid == 0) {
        u64 const hs = d_huf_write_ctable(sh.header, cap < sizeof(sh.header) ? cap : sizeof(sh.header), sh.ctable, msv, huffLog, sh.wksp);
        if (is_err(hs)) { sh.flag = 1; sh.verdict = hs; }
        else if (hs + 12 >= n) { sh.flag = 1; sh.verdict = 0; }
        else sh.hSize = (u32)hs;
    }
    __syncthreads();
    if (sh.flag) { if (tid == 0) csizes[b] = sh.verdict; return; }
    u32 const hSize = sh.hSize;
    u64 const capLeft = cap - hSize;
    if (capLeft < 6 + 1 + 1 + 1 + 8 || n < 12) { if (tid == 0) csizes[b] = 0; return; }              // :564-565

    // ---- pass A: bit length of every thread's run; stream k is owned by threads 64k..64k+63 ----
    u32 const seg = (n + 3) / 4;
    int const k = tid >> 6, j = tid & 63;
    u32 const segBeg = (u32)k * seg;
    u32 const segEnd = (k < 3) ? segBeg + seg : n;
    u32 const segLen = segEnd - segBeg;
    u32 const run = (segLen + 63) / 64;
    u32 const offHi = min((u32)j * run, segLen), offLo = min((u32)(j + 1) * run, segLen);
    u32 const hiC = segEnd - offHi, loIdx = segEnd - offLo;      // this thread's run = symbols [loIdx, hiC), emitted high to low
    {   u32 bits = 0;
        for (u32 i = loIdx; i < hiC; i++) bits += sh.ctable[sp[i]] >> 16;
        sh.chunkBits[tid] = bits;
    }
    __syncthreads();
    if (j == 0) {                                  // exclusive scan over the 64 runs of this stream
        u32 acc = 0;
        for (int t = 0; t < 64; t++) { sh.chunkOff[64 * k + t] = acc; acc += sh.chunkBits[64 * k + t]; }
        sh.streamBytes[k] = acc;                   // bits for now
    }
    __syncthreads();
    if (tid == 0) {                                // sizes with the writer's capacity rule, stream after stream (:566-600, bitstream.h:190,246,258)
        u64 op = 6; bool fits = true;
        for (int t = 0; t < 4 && fits; t++) {
            u64 const capk = capLeft - op;
            u64 const tot = (u64)sh.streamBytes[t] + 1;                       // + end mark
            if (capk < 8 || capk <= 8 || (tot >> 3) >= capk - 8) { fits = false; break; }
            sh.streamOff[t] = (u32)op;
            sh.streamBytes[t] = (u32)((tot + 7) >> 3);
            op += sh.streamBytes[t];
        }
        u64 const total = hSize + op;
        if (!fits) { sh.flag = 1; sh.verdict = 0; }
        else if (total >= (u64)n - 1) { sh.flag = 1; sh.verdict = 0; }                                // :625
        else sh.total = (u32)total;
    }
    __syncthreads();
    if (sh.flag) { if (tid == 0) csizes[b] = sh.verdict; return; }
    u32 const total = sh.total;

    // ---- pass B: build the block image in shared memory ----
    u32 const al = (u32)(reinterpret_cast<u64>(d) & 15);          // image byte i <-> d[i - al]
    u32 const imgWords = (al + total + 3) / 4 + 1;
    for (u32 i = tid; i < imgWords; i += THREADS) image[i] = 0;
    __syncthreads();
    {   u8* const img8 = reinterpret_cast<u8*>(image);
        for (u32 i = tid; i < hSize; i += THREADS) img8[al + i] = sh.header[i];
        if (tid < 3) { u32 const v = sh.streamBytes[tid]; img8[al + hSize + 2 * tid] = (u8)v; img8[al + hSize + 2 * tid + 1] = (u8)(v >> 8); }
    }
    __syncthreads();
    {   u64 const P = 8ull * (al + hSize + sh.streamOff[k]) + sh.chunkOff[tid];
        u32* wp = image + (P >> 5);
        unsigned held = (unsigned)(P & 31);
        u64 acc = 0;
        for (u32 i = hiC; i-- > loIdx;) {
            u32 const e = sh.ctable[sp[i]];
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
        u32 const first16 = (al + 15) & ~15u;                       // first 16-aligned image offset with data
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
    u32 const stageBytes = g.blockSize <= 65536 ? ((g.blockSize + 15u + 16u) & ~15u) : 0u;    // 2 x 128 KB would not fit in 227 KB
    size_t const imageBytes = (size_t)g.blockSize + 64 + 16;                       // accepted blocks are < n bytes
    size_t const histBytes = hufe::WARPS * 256 * sizeof(u32);
    size_t const smem = ((sizeof(hufe::Shared) + 15) & ~(size_t)15) + stageBytes + (imageBytes > histBytes ? imageBytes : histBytes) + 16;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(hufe::huf_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    hufe::huf_encode_kernel<<<g.nBlocks, hufe::THREADS, smem, stream>>>(g, (u8*)cbuf, csizes, (const u8*)src, msv, tlog, stageBytes);
    return cudaGetLastError();
}

}  // namespace fseb
