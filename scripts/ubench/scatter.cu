// scatter.cu -- micro-benchmark of the decoder's global access pattern: every lane walks its OWN region (regions 8 KB / 4.3 KB
// apart inside 32 KB blocks, 32 lanes of a warp in 32 different blocks), 16 or 32 bytes per access.  Answers: what does the
// per-lane-scattered 16-byte store / load cost at 1 GiB, and what do 32-byte (v8) accesses or warp-coalesced lines buy?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned u32;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// mode 0: STG.128 per lane per step; 1: st.v8 (32 B) per lane per 2 steps; 2: warp-coalesced (lane-transposed) 128 B lines
template <int MODE>
__global__ void __launch_bounds__(256, 4) store_k(unsigned char* out, u32 nBlocks, u32 gEff)
{
    int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int const col = 2 * lane + (warp >> 2), strm = warp & 3;
    u32 const b = blockIdx.x * gEff + col;
    if (col >= (int)gEff || b >= nBlocks) return;
    unsigned char* p = out + (size_t)b * 32768 + strm * 8192;
    u32 v = tid;
    if (MODE == 0) for (int i = 0; i < 8192; i += 16) { v = v * 1664525u + 1013904223u; *reinterpret_cast<uint4*>(p + i) = make_uint4(v, v + 1, v + 2, v + 3); }
    if (MODE == 1) for (int i = 0; i < 8192; i += 32) {
        v = v * 1664525u + 1013904223u;
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p + i), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3), "r"(v + 4), "r"(v + 5), "r"(v + 6), "r"(v + 7) : "memory");
    }
}
template <int MODE>
__global__ void __launch_bounds__(256, 4) load_k(const unsigned char* in, u32 nBlocks, u32 gEff, u32* sink)
{
    int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int const col = 2 * lane + (warp >> 2), strm = warp & 3;
    u32 const b = blockIdx.x * gEff + col;
    if (col >= (int)gEff || b >= nBlocks) return;
    const unsigned char* p = in + (size_t)b * 17280 + strm * 4320;       // compressed-size stride (P14: ~17.2 KB per block)
    u32 acc = 0;
    if (MODE == 0) for (int i = 4304; i >= 0; i -= 16) { uint4 x = __ldg(reinterpret_cast<const uint4*>(p + i)); acc += x.x ^ x.y ^ x.z ^ x.w; }
    if (MODE == 1) for (int i = 4288; i >= 0; i -= 32) {
        u32 a0, a1, a2, a3, a4, a5, a6, a7;
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(a4), "=r"(a5), "=r"(a6), "=r"(a7) : "l"(p + i));
        acc += a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}
// coalesced reference: the same bytes moved with fully coalesced 16-byte accesses (the roofline-style copy)
__global__ void __launch_bounds__(256, 4) copy_k(uint4* out, const uint4* in, size_t n16)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
template <typename F> float timeit(F f, int reps = 5)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < reps; r++) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = ms < best ? ms : best; }
    return best;
}
int main()
{
    u32 const nBlocks = 32768, gEff = 56; u32 const grid = (nBlocks + gEff - 1) / gEff;
    unsigned char *out, *in; u32* sink;
    CK(cudaMalloc(&out, (size_t)nBlocks * 32768)); CK(cudaMalloc(&in, (size_t)nBlocks * 17280 + 4096)); CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(in, 1, (size_t)nBlocks * 17280 + 4096));
    size_t const dynSmem = 55 * 1024;   // same shared-memory footprint as the decoder: four CTAs per SM, a minimal L1
    CK(cudaFuncSetAttribute(store_k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynSmem)); CK(cudaFuncSetAttribute(store_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynSmem));
    CK(cudaFuncSetAttribute(load_k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynSmem)); CK(cudaFuncSetAttribute(load_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynSmem));
    for (size_t sm : { (size_t)0, dynSmem }) {
        printf("dynamic smem per CTA = %zu\n", sm);
        printf("  store 16 B/lane   : %.3f ms per GiB\n", timeit([&] { store_k<0><<<grid, 256, sm>>>(out, nBlocks, gEff); }));
        printf("  store 32 B/lane   : %.3f ms per GiB\n", timeit([&] { store_k<1><<<grid, 256, sm>>>(out, nBlocks, gEff); }));
        printf("  load  16 B/lane   : %.3f ms per 0.527 GiB\n", timeit([&] { load_k<0><<<grid, 256, sm>>>(in, nBlocks, gEff, sink); }));
        printf("  load  32 B/lane   : %.3f ms per 0.527 GiB\n", timeit([&] { load_k<1><<<grid, 256, sm>>>(in, nBlocks, gEff, sink); }));
    }
    printf("  coalesced copy 1 GiB: %.3f ms\n", timeit([&] { copy_k<<<148 * 8, 256>>>((uint4*)out, (const uint4*)out + (size_t)nBlocks * 1024, (size_t)nBlocks * 1024); }));
    return 0;
}
