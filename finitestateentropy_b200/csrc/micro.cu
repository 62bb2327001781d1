// micro.cu -- single-CTA kernels behind the reference's table-level entry points (lib/fse.h:135-247,
// lib/hist.h:30-75, lib/huf.h:188-272 static section).  They run the very same device routines the
// batched kernels use, on one table, so that HIST_count / FSE_normalizeCount / FSE_buildCTable /
// FSE_buildDTable / HUF_buildCTable / HUF_readDTableX1 ... are honest drop-ins executing on the GPU.
#include "common.cuh"
#include "fse_dev.cuh"
#include "bitsrc_dev.cuh"
#include "sink_dev.cuh"
#include "huf_dev.cuh"
#include "huf_build_dev.cuh"
#include "micro.h"

namespace fseb {

// ---- HIST_count (lib/hist.c:163-180): any size, one CTA ----
__global__ void __launch_bounds__(256) hist_kernel(const u8* src, u64 n, u32 declared, u32* out /*256 counts, then msv, then pad*/, u64* ret)
{
    __shared__ u32 h[8][256];
    __shared__ u32 tot[256];
    int const tid = threadIdx.x;
    for (int i = tid; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
    __syncthreads();
    u32* const mine = h[tid >> 5];
    for (u64 i = tid; i < n; i += 256) atomicAdd(&mine[src[i]], 1u);
    __syncthreads();
    {   u32 c = 0; for (int w = 0; w < 8; w++) c += h[w][tid]; tot[tid] = c; }
    __syncthreads();
    if (tid == 0) {
        if (declared > 255) declared = 255;
        if (n == 0) { for (u32 i = 0; i <= declared; i++) out[i] = 0; out[256] = 0; *ret = 0; return; }
        u32 top = 255; while (!tot[top]) top--;
        if (declared < 255 && top > declared) { *ret = err(E_MSV_TOO_SMALL); return; }
        u32 best = 0;
        for (u32 i = 0; i < 256; i++) best = tot[i] > best ? tot[i] : best;
        for (u32 i = 0; i <= declared; i++) out[i] = tot[i];
        out[256] = top; *ret = best;
    }
}

// ---- everything O(alphabet)/O(table): one kernel, op-code dispatched ----
// buf layout is op specific (see capi.cu); a[] carries scalars.
__global__ void __launch_bounds__(256) micro_kernel(int op, MicroArgs A, u8* buf, u64* ret)
{
    __shared__ u32 s_count[256];
    __shared__ u32 s_ct[256];
    __shared__ HNode s_nodes[2 * 256 + 2];
    __shared__ u32 s_a[256], s_b[256];
    __shared__ u32 s_w[384];
    int const tid = threadIdx.x;
    switch (op) {
    case MOP_NORMALIZE: {          // in: count u32[msv+1] @0 ; out: norm i16[msv+1] @4096
        if (tid == 0) *ret = d_normalize((short*)(buf + 4096), (unsigned)A.a[0], (const unsigned*)buf, A.a[1], (unsigned)A.a[2]);
        break; }
    case MOP_WRITE_NCOUNT: {       // in: norm @0 ; out: bytes @4096
        if (tid == 0) *ret = d_write_ncount(buf + 4096, A.a[0], (const short*)buf, (unsigned)A.a[1], (unsigned)A.a[2]);
        break; }
    case MOP_READ_NCOUNT: {        // in: header bytes @0 (size a0), msv in a1 ; out: norm @4096, msv/tl as u32 @8192
        if (tid == 0) {
            unsigned msv = (unsigned)A.a[1], tl = 0;
            *ret = d_read_ncount((short*)(buf + 4096), &msv, &tl, buf, A.a[0]);
            ((u32*)(buf + 8192))[0] = msv; ((u32*)(buf + 8192))[1] = tl;
        }
        break; }
    case MOP_BUILD_CTABLE: {       // in: norm @0 ; out: ct @4096 ; scratch @65536
        if (tid == 0) {
            unsigned const msv = (unsigned)A.a[0], tl = (unsigned)A.a[1];
            if (tl > 13 || msv > 4095) { *ret = err(E_TLOG_TOO_LARGE); break; }
            d_build_ctable_serial((u32*)(buf + 4096), (const short*)buf, msv, tl, (u16*)(buf + 65536), (u32*)(buf + 65536 + 32768));
            *ret = 0;
        }
        break; }
    case MOP_BUILD_DTABLE: {       // in: norm @0 ; out: dt @4096 ; a2 = wide
        if (tid == 0) {
            unsigned const msv = (unsigned)A.a[0], tl = (unsigned)A.a[1];
            if (A.a[2]) *ret = d_build_dtable_serial<true>((u32*)(buf + 4096), (const short*)buf, msv, tl, U16_MAX_SV, U16_MAX_TLOG, (u16*)(buf + 65536), (u16*)(buf + 65536 + 32768));
            else *ret = d_build_dtable_serial<false>((u32*)(buf + 4096), (const short*)buf, msv, tl, FSE_MAX_SV, FSE_MAX_TLOG, (u16*)(buf + 65536), (u16*)(buf + 65536 + 32768));
        }
        break; }
    case MOP_HUF_BUILD_CTABLE: {   // in: count u32[256] @0 ; out: ctable u32[256] @4096
        s_count[tid] = ((u32)tid <= A.a[0]) ? ((const u32*)buf)[tid] : 0;
        __syncthreads();
        u64 const r = cta_huf_build_ctable(s_ct, s_count, (u32)A.a[0], (u32)A.a[1], s_nodes, s_a, s_b);
        if (!is_err(r)) ((u32*)(buf + 4096))[tid] = s_ct[tid];
        if (tid == 0) *ret = r;
        break; }
    case MOP_HUF_WRITE_CTABLE: {   // in: ctable u32[256] @0 ; out: bytes @4096 ; a0 = cap, a1 = msv, a2 = huffLog
        if (tid == 0) *ret = d_huf_write_ctable(buf + 4096, A.a[0] < 136 ? A.a[0] : 136, (const u32*)buf, (unsigned)A.a[1], (unsigned)A.a[2], s_w);
        break; }
    case MOP_HUF_READ_STATS: {     // in: bytes @0 (a0) ; out: weights @4096 (hwSize a1), rankStats u32[13] @8192, nbSym,tl @8192+64
        if (tid == 0) {
            u32 nb = 0, tl = 0;
            *ret = d_huf_read_stats(buf + 4096, A.a[1], (u32*)(buf + 8192), &nb, &tl, buf, A.a[0]);
            ((u32*)(buf + 8192 + 64))[0] = nb; ((u32*)(buf + 8192 + 64))[1] = tl;
        }
        break; }
    case MOP_HUF_READ_DTABLE_X1: { // in: bytes @0 (a0), a1 = DTable header word ; out: dtable u32[1+2048] @16384
        if (tid == 0) {
            u8* const weights = buf + 4096;
            u32 rank[17]; u32 nb = 0, tl = 0;
            u32* const dt = (u32*)(buf + 16384);
            u16* const cells = (u16*)(dt + 1);
            u64 const h = d_huf_read_stats(weights, 256, rank, &nb, &tl, buf, A.a[0]);
            u32 const hdr = (u32)A.a[1];
            dt[0] = hdr;
            if (is_err(h)) { *ret = h; break; }
            if (tl > (hdr & 0xFF) + 1) { *ret = err(E_TLOG_TOO_LARGE); break; }      // huf_decompress.c:143
            dt[0] = (hdr & 0xFF0000FFu) | (tl << 16);
            u32 next = 0;
            for (u32 n = 1; n < tl + 1; n++) { u32 const cur = next; next += rank[n] << (n - 1); rank[n] = cur; }
            for (u32 n = 0; n < nb; n++) {
                u32 const w = weights[n], span = (1u << w) >> 1, cell = n | ((tl + 1 - w) << 8);
                for (u32 u = 0; u < span; u++) cells[rank[w] + u] = (u16)cell;
                rank[w] += span;
            }
            *ret = h;
        }
        break; }
    default: if (tid == 0) *ret = err(E_GENERIC);
    }
}

cudaError_t launch_hist(const void* src, u64 n, u32 declared, u32* out, u64* ret, cudaStream_t s)
{ hist_kernel<<<1, 256, 0, s>>>((const u8*)src, n, declared, out, ret); return cudaGetLastError(); }
cudaError_t launch_micro(int op, const MicroArgs& A, void* buf, u64* ret, cudaStream_t s)
{ micro_kernel<<<1, 256, 0, s>>>(op, A, (u8*)buf, ret); return cudaGetLastError(); }

}  // namespace fseb
