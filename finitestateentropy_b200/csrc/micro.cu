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
#include "huf_x2_dev.cuh"
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

// ---- FSE_countU16 (lib/fseU16.c:121-145): histogram of 16-bit symbols, one CTA, counters in global scratch ----
// out: count u32[declared+1], then out[declared+1] = largest present symbol.  ret: largest count, 0 for an empty input, or
// maxSymbolValue_tooSmall if a symbol exceeds `declared` (the reference stops at the first such symbol; the counts it leaves
// behind are unspecified, ours are zeros).
__global__ void __launch_bounds__(256) hist16_kernel(const u16* src, u64 n, u32 declared, u32* out, u64* ret)
{
    __shared__ u32 s_bad, s_top, s_best;
    int const tid = threadIdx.x;
    for (u32 i = tid; i <= declared + 1; i += 256) out[i] = 0;
    if (tid == 0) { s_bad = 0; s_top = 0; s_best = 0; }
    __syncthreads();
    u32 bad = 0;
    for (u64 i = tid; i < n; i += 256) { u32 const v = src[i]; if (v > declared) bad = 1; else atomicAdd(&out[v], 1u); }
    if (bad) s_bad = 1;
    __syncthreads();
    if (s_bad) { for (u32 i = tid; i <= declared; i += 256) out[i] = 0; if (tid == 0) *ret = err(E_MSV_TOO_SMALL); return; }
    u32 top = 0, best = 0;
    for (u32 i = tid; i <= declared; i += 256) { u32 const c = out[i]; if (c) { top = i; best = c > best ? c : best; } }
    atomicMax(&s_top, top); atomicMax(&s_best, best);
    __syncthreads();
    if (tid == 0) { out[declared + 1] = n ? s_top : 0; *ret = n ? s_best : 0; }
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
    // ---- payload coding with a caller-supplied table image (the *_usingCTable / *_usingDTable entry points): table @0,
    //      input @a2 (a0 bytes), output @a3 (capacity a1).  One lane per stream; these serve single calls, not throughput.
    case MOP_FSE_ENCODE_CT: {      // FSE_compress_usingCTable (fse_compress.c:554-623)
        if (tid == 0) *ret = d_fse_encode_serial(buf + A.a[3], A.a[1], buf + A.a[2], A.a[0], (const u32*)buf);
        break; }
    case MOP_FSE_DECODE_DT: {      // FSE_decompress_usingDTable (fse_decompress.c:178-252)
        if (tid == 0) *ret = d_fse_decode_serial(buf + A.a[3], A.a[1], buf + A.a[2], A.a[0], (const u32*)buf);
        break; }
    case MOP_HUF_ENCODE4X_CT: {    // HUF_compress4X_usingCTable (huf_compress.c:552-603): lane k codes segment k, lane 0 assembles
        __shared__ u64 s_len[4];
        const u32* const ct = (const u32*)buf;
        const u8* const in = buf + A.a[2]; u8* const out = buf + A.a[3];
        u64 const n = A.a[0], cap = A.a[1];
        u64 const seg = (n + 3) / 4;
        bool const refuse = cap < 6 + 1 + 1 + 1 + 8 || n < 12;                       // :564-565
        u8* const stage = buf + A.a[3] + ((cap + 15) & ~15ull);                     // 4 private areas of `cap` bytes behind the output
        if (tid < 4 && !refuse) {
            u64 const beg = seg * tid, end = tid < 3 ? seg * (tid + 1) : n;
            BitSink sk; sink_open(sk, stage + tid * ((cap + 15) & ~15ull), cap);     // capacity is judged by lane 0 with the real remaining room
            for (u64 i = end; i-- > beg;) { u32 const e = ct[in[i]]; sink_put(sk, e & 0xFFFF, e >> 16); }
            sink_put(sk, 1, 1);                                                      // end mark
            s_len[tid] = sk.nbits;
            if (sk.held) { if (sk.nbytes < cap) sk.out[sk.nbytes] = (u8)sk.acc; }
        }
        __syncthreads();
        if (tid == 0) {
            if (refuse) { *ret = 0; break; }
            u64 op = 6; u64 r = 0; bool ok = true;
            for (int k = 0; k < 4 && ok; k++) {
                u64 const room = cap - op;                                           // BIT_initCStream / closeCStream rules (bitstream.h:183-260)
                u64 const bits = s_len[k], bytes = (bits + 7) >> 3;
                if (room <= 8 || (bits >> 3) >= room - 8) { ok = false; break; }
                const u8* const sp = stage + k * ((cap + 15) & ~15ull);
                for (u64 i = 0; i < bytes; i++) out[op + i] = sp[i];
                if (k < 3) { out[2 * k] = (u8)bytes; out[2 * k + 1] = (u8)(bytes >> 8); }
                op += bytes;
            }
            r = ok ? op : 0;
            *ret = r;
        }
        break; }
    case MOP_HUF_DECODE4X1_DT: {   // HUF_decompress4X1_usingDTable (huf_decompress.c:262-354): lane k decodes stream k into segment k
        __shared__ u64 s_init[4]; __shared__ u32 s_done[4];
        const u32* const dtab = (const u32*)buf;
        const u16* const cells = (const u16*)(dtab + 1);                             // HUF_DEltX1 { byte, nbBits }
        u32 const dtLog = (dtab[0] >> 16) & 0xFF;
        const u8* const c = buf + A.a[2]; u8* const out = buf + A.a[3];
        u64 const cs = A.a[0], n = A.a[1];
        bool bad = cs < 10;                                                          // :268
        u64 l1 = 0, l2 = 0, l3 = 0, l4 = 0;
        if (!bad) {
            l1 = c[0] | ((u64)c[1] << 8); l2 = c[2] | ((u64)c[3] << 8); l3 = c[4] | ((u64)c[5] << 8);
            if (l1 + l2 + l3 + 6 > cs) bad = true; else l4 = cs - (l1 + l2 + l3 + 6);   // the reference would read out of bounds here
        }
        u64 const seg = (n + 3) / 4;
        if (!bad && 3 * seg > n) bad = true;                                         // dstSize < 6: the reference writes out of bounds (documented deviation)
        if (tid < 4) {
            u64 ie = 0; u32 done = 0;
            if (!bad) {
                u64 const lens[4] = { l1, l2, l3, l4 };
                u64 off = 6; for (int k = 0; k < tid; k++) off += lens[k];
                long long p = (long long)(seg * tid); long long const pe = tid < 3 ? (long long)(seg * (tid + 1)) : (long long)n;
                BitSrc b;
                ie = bs_open(b, c + off, lens[tid]);
                if (!is_err(ie)) {
                    ie = 0;
                    auto sym = [&]() { u32 const cell = cells[bs_peek_fast(b, dtLog)]; b.used += cell >> 8; out[p++] = (u8)cell; };
                    while ((bs_refill(b) == SRC_MORE) & (p < pe - 3)) { sym(); sym(); sym(); sym(); }   // HUF_decodeStreamX1 :214-237
                    while (p < pe) sym();
                    done = bs_exhausted(b) ? 1u : 0u;                                // :348-349
                }
            }
            s_init[tid] = ie; s_done[tid] = done;
        }
        __syncthreads();
        if (tid == 0) {
            u64 r = n;
            if (bad) r = err(E_CORRUPT);
            else {
                bool initFailed = false;
                for (int k = 0; k < 4; k++) if (is_err(s_init[k])) { r = s_init[k]; initFailed = true; break; }   // CHECK_F in stream order (:297-300)
                if (!initFailed && !(s_done[0] & s_done[1] & s_done[2] & s_done[3])) r = err(E_CORRUPT);
            }
            *ret = r;
        }
        break; }
    case MOP_HUF_ENCODE1X_CT: {    // HUF_compress1X_usingCTable (huf_compress.c:457-502): one stream, last symbol first
        if (tid == 0) {
            const u32* const ct = (const u32*)buf; const u8* const in = buf + A.a[2];
            u64 const n = A.a[0], cap = A.a[1];
            if (cap < 8) { *ret = 0; break; }                                        // :470
            BitSink sk; sink_open(sk, buf + A.a[3], cap);
            if (!sk.usable) { *ret = 0; break; }
            for (u64 i = n; i-- > 0;) { u32 const e = ct[in[i]]; sink_put(sk, e & 0xFFFF, e >> 16); }
            *ret = sink_close(sk);
        }
        break; }
    case MOP_HUF_DECODE1X1_DT: {   // HUF_decompress1X1_usingDTable (huf_decompress.c:240-260)
        if (tid == 0) {
            const u32* const dtab = (const u32*)buf;
            const u16* const cells = (const u16*)(dtab + 1);
            u32 const dtLog = (dtab[0] >> 16) & 0xFF;
            u8* const out = buf + A.a[3];
            long long p = 0; long long const pe = (long long)A.a[1];
            BitSrc b;
            u64 const e = bs_open(b, buf + A.a[2], A.a[0]);
            if (is_err(e)) { *ret = e; break; }
            auto sym = [&]() { u32 const cell = cells[bs_peek_fast(b, dtLog)]; b.used += cell >> 8; out[p++] = (u8)cell; };
            while ((bs_refill(b) == SRC_MORE) & (p < pe - 3)) { sym(); sym(); sym(); sym(); }
            while (p < pe) sym();
            *ret = bs_exhausted(b) ? (u64)pe : err(E_CORRUPT);
        }
        break; }
    case MOP_HUF_READ_DTABLE_X2: { // HUF_readDTableX2 (huf_decompress.c:551-649): in: bytes @0 (a0), a1 = DTable header word ; out: dtable u32[1+4096] @32768
        if (tid == 0) *ret = d_huf_build_dtable_x2((u32*)(buf + 32768), (u32)A.a[1], buf + 4096, buf + 8192, buf + 8192 + 256, buf, A.a[0]);
        break; }
    case MOP_HUF_DECODE4X2_DT:     // HUF_decompress4X2_usingDTable (huf_decompress.c:749-862): lane k decodes stream k with the double-symbol table
    case MOP_HUF_DECODE1X2_DT: {   // HUF_decompress1X2_usingDTable (:722-747): one stream
        __shared__ u64 s_init2[4]; __shared__ u32 s_done2[4];
        u64 const r = cta_huf_decode_x2(op == MOP_HUF_DECODE4X2_DT, (const u32*)buf, buf + A.a[2], A.a[0], buf + A.a[3], A.a[1], s_init2, s_done2);
        if (tid == 0) *ret = r;
        break; }
    default: if (tid == 0) *ret = err(E_GENERIC);
    }
}

cudaError_t launch_hist(const void* src, u64 n, u32 declared, u32* out, u64* ret, cudaStream_t s)
{ hist_kernel<<<1, 256, 0, s>>>((const u8*)src, n, declared, out, ret); return cudaGetLastError(); }
cudaError_t launch_hist16(const void* src, u64 n, u32 declared, u32* out, u64* ret, cudaStream_t s)
{ hist16_kernel<<<1, 256, 0, s>>>((const u16*)src, n, declared, out, ret); return cudaGetLastError(); }
cudaError_t launch_micro(int op, const MicroArgs& A, void* buf, u64* ret, cudaStream_t s)
{ micro_kernel<<<1, 256, 0, s>>>(op, A, (u8*)buf, ret); return cudaGetLastError(); }

}  // namespace fseb
