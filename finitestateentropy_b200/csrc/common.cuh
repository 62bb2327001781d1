// common.cuh -- shared device/host helpers for the B200 (sm_100a) entropy-coding kernels.
//
// Error convention: identical to the reference's (lib/error_private.h:77-79,
// lib/error_public.h:45-56): every entry point returns size_t, errors are (size_t)-code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace fseb {

enum : unsigned {
    E_OK = 0, E_GENERIC = 1, E_DST_TOO_SMALL = 2, E_SRC_WRONG = 3, E_CORRUPT = 4,
    E_TLOG_TOO_LARGE = 5, E_MSV_TOO_LARGE = 6, E_MSV_TOO_SMALL = 7, E_WKSP_TOO_SMALL = 8, E_MAXCODE = 9
};

typedef unsigned long long u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

__host__ __device__ __forceinline__ u64 err(unsigned code) { return (u64)0 - (u64)code; }
__host__ __device__ __forceinline__ bool is_err(u64 r) { return r > err(E_MAXCODE); }

// constants fixed by the wire format (lib/fse.h:641-676, lib/huf.h:117-119,72)
constexpr unsigned FSE_MIN_TLOG = 5, FSE_MAX_TLOG = 12, FSE_DEF_TLOG = 11, FSE_ABS_TLOG = 15, FSE_MAX_SV = 255;
constexpr unsigned HUF_MAX_TLOG = 12, HUF_DEF_TLOG = 11, HUF_MAX_SV = 255, HUF_BLOCK_MAX = 128 * 1024;
constexpr unsigned U16_MAX_SV = 286, U16_MAX_TLOG = 13, U16_DEF_TLOG = 12;

// Transient per-block result of the batch Huff0 decoder: "rejected by the single-symbol end-of-stream rule, to be re-examined
// under the double-symbol decoder's rules" (huf_x2_fixup.cu replaces it before the call returns).  Neither a size nor an error code.
constexpr u64 HUF_X2_PENDING = 0x8000000000000004ull;

__device__ __forceinline__ unsigned hibit(unsigned v) { return 31u - (unsigned)__clz((int)v); }   // v != 0
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// unaligned little-endian reads through a byte pointer (generic address space)
__device__ __forceinline__ u32 rd16(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8); }
__device__ __forceinline__ u32 rd32(const u8* p) { return rd16(p) | (rd16(p + 2) << 16); }

// -------------------------------------------------------------------------------------------
// Uniform batch geometry: the split programs/bench.c:530-548 performs on a flat buffer.
//   block b covers [b*blockSize, min(total,(b+1)*blockSize)); compressed block b lives in the
//   fixed slot cbuf + b*slot (slot = FSE_compressBound(blockSize) in the reference harness).
// -------------------------------------------------------------------------------------------
struct BatchGeom {
    u64 total;       // uncompressed bytes in the whole batch
    u32 blockSize;   // uncompressed bytes per block (last block may be shorter)
    u32 slot;        // byte stride between compressed blocks == per-block dst capacity
    u32 nBlocks;
};
__host__ __device__ __forceinline__ u32 block_len(const BatchGeom& g, u32 b)
{
    u64 const off = (u64)b * g.blockSize;
    u64 const left = g.total - off;
    return left < g.blockSize ? (u32)left : g.blockSize;
}

}  // namespace fseb
