// huf_build_dev.cuh -- device-side Huffman code construction and tree-header writer.
//
//   HUF_buildCTable_wksp  lib/huf_compress.c:338-410  (HUF_sort :307-329, HUF_setMaxHeight :215-291)
//   HUF_writeCTable       lib/huf_compress.c:114-147  (HUF_compressWeights :63-103)
// Tie-breaks that decide the compressed bytes are kept: sort = decreasing count, ties by increasing
// symbol; the merge takes the internal node on equal counts; the depth limiter's repayment order;
// codes are canonical with the longest lengths numbered from 0, symbol order inside a length.
// CTable cell = val | nbBits << 16 (the {U16 val; BYTE nbBits} layout of huf_compress.c:106-109).
#pragma once
#include "common.cuh"
#include "fse_dev.cuh"
#include "sink_dev.cuh"

namespace fseb {

struct HNode { u32 count; u16 parent; u8 sym; u8 len; };

// one lane; nd[-1] must be addressable (sentinel)
__device__ inline u32 d_huf_limit_depth(HNode* nd, u32 last, u32 maxBits)
{
    u32 const deepest = nd[last].len;
    if (deepest <= maxBits) return deepest;
    int debt = 0;
    u32 const unit = 1u << (deepest - maxBits);
    int n = (int)last;
    u32 const NONE = 0xF0F0F0F0u;
    u32 lastOfRank[HUF_MAX_TLOG + 2];
    while (nd[n].len > maxBits) {
        debt += (int)(unit - (1u << (deepest - nd[n].len)));
        nd[n].len = (u8)maxBits; n--;
    }
    while (nd[n].len == maxBits) n--;
    debt >>= (deepest - maxBits);
    for (u32 i = 0; i < HUF_MAX_TLOG + 2; i++) lastOfRank[i] = NONE;
    {   u32 cur = maxBits;
        for (int pos = n; pos >= 0; pos--) {
            if (nd[pos].len >= cur) continue;
            cur = nd[pos].len;
            lastOfRank[maxBits - cur] = (u32)pos;
        }
    }
    while (debt > 0) {
        u32 dec = hibit((u32)debt) + 1;
        for (; dec > 1; dec--) {
            u32 const hi = lastOfRank[dec], lo = lastOfRank[dec - 1];
            if (hi == NONE) continue;
            if (lo == NONE) break;
            if (nd[hi].count <= 2 * nd[lo].count) break;
        }
        while (dec <= HUF_MAX_TLOG && lastOfRank[dec] == NONE) dec++;
        debt -= 1 << (dec - 1);
        if (lastOfRank[dec - 1] == NONE) lastOfRank[dec - 1] = lastOfRank[dec];
        nd[lastOfRank[dec]].len++;
        if (lastOfRank[dec] == 0) lastOfRank[dec] = NONE;
        else {
            lastOfRank[dec]--;
            if (nd[lastOfRank[dec]].len != maxBits - dec) lastOfRank[dec] = NONE;
        }
    }
    while (debt < 0) {
        if (lastOfRank[1] == NONE) {
            while (nd[n].len == maxBits) n--;
            nd[n + 1].len--;
            lastOfRank[1] = (u32)(n + 1);
            debt++;
            continue;
        }
        nd[lastOfRank[1] + 1].len--;
        lastOfRank[1]++;
        debt++;
    }
    return maxBits;
}

// CTA-cooperative (blockDim.x == 256).  count/ctable/lenOf/firstVal live in shared memory.
// store: HNode[2*256+2].  Returns the maximum code length, or an error code; contains barriers.
__device__ inline u64 cta_huf_build_ctable(u32* ctable, const u32* count, u32 msv, u32 maxBits,
                                           HNode* store, u32* lenOf, u32* firstVal)
{
    int const tid = threadIdx.x;
    HNode* const nd = store + 1;
    __shared__ u32 s_ret;
    if (!maxBits) maxBits = HUF_DEF_TLOG;
    if (msv > HUF_MAX_SV) return err(E_MSV_TOO_LARGE);
    for (int i = tid; i < 2 * 256 + 2; i += blockDim.x) { HNode z; z.count = 0; z.parent = 0; z.sym = 0; z.len = 0; store[i] = z; }
    __syncthreads();
    if ((u32)tid <= msv) {                                   // rank by counting == stable sort by decreasing count
        u32 const c = count[tid];
        u32 rank = 0;
        for (u32 jj = 0; jj <= msv; jj++) { u32 const cj = count[jj]; rank += (cj > c) | ((cj == c) & (jj < (u32)tid)); }
        nd[rank].count = c; nd[rank].sym = (u8)tid;
    }
    __syncthreads();
    if (tid == 0) {
        int last = (int)msv; while (nd[last].count == 0) last--;
        int fresh = 256, leaf = last, root = fresh + leaf - 1, inner = fresh;
        nd[fresh].count = nd[leaf].count + nd[leaf - 1].count;
        nd[leaf].parent = nd[leaf - 1].parent = (u16)fresh;
        fresh++; leaf -= 2;
        for (int n = fresh; n <= root; n++) nd[n].count = 1u << 30;
        nd[-1].count = 1u << 31;
        while (fresh <= root) {
            int const a = (nd[leaf].count < nd[inner].count) ? leaf-- : inner++;
            int const b = (nd[leaf].count < nd[inner].count) ? leaf-- : inner++;
            nd[fresh].count = nd[a].count + nd[b].count;
            nd[a].parent = nd[b].parent = (u16)fresh;
            fresh++;
        }
        nd[root].len = 0;
        for (int n = root - 1; n >= 256; n--) nd[n].len = (u8)(nd[nd[n].parent].len + 1);
        for (int n = 0; n <= last; n++) nd[n].len = (u8)(nd[nd[n].parent].len + 1);
        u32 const mb = d_huf_limit_depth(nd, (u32)last, maxBits);
        s_ret = mb;
        if (mb <= HUF_MAX_TLOG) {
            u32 perLen[HUF_MAX_TLOG + 1];
            for (u32 i = 0; i <= HUF_MAX_TLOG; i++) perLen[i] = 0;
            for (int n = 0; n <= last; n++) perLen[nd[n].len]++;
            u32 v = 0;
            for (u32 i = 0; i <= HUF_MAX_TLOG; i++) firstVal[i] = 0;
            for (int n = (int)mb; n > 0; n--) { firstVal[n] = v; v = (v + perLen[n]) >> 1; }
        }
    }
    __syncthreads();
    u32 const mb = s_ret;
    if (mb > HUF_MAX_TLOG) return err(E_GENERIC);
    if ((u32)tid <= msv) lenOf[nd[tid].sym] = nd[tid].len;
    __syncthreads();
    if ((u32)tid <= msv) {                                   // value = first value of the length + number of earlier symbols of that length
        u32 const len = lenOf[tid];
        u32 before = 0;
        for (u32 jj = 0; jj < (u32)tid; jj++) before += (lenOf[jj] == len);
        ctable[tid] = ((firstVal[len] + before) & 0xFFFF) | (len << 16);
    } else ctable[tid & 255] = 0;
    __syncthreads();
    return mb;
}

// 256 keys, element e = i * 32 + lane, sorted into DEcreasing order (bitonic network: shuffles for partners in other
// lanes, plain compare-exchange for partners in the same lane).  Keys must be distinct.
__device__ __forceinline__ void warp_sort256_desc(u32 (&key)[8], unsigned lane)
{
    #pragma unroll
    for (u32 k = 2; k <= 256; k <<= 1) {
        #pragma unroll
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                u32 const dj = j >> 5;
                #pragma unroll
                for (u32 i = 0; i < 8; i++) {
                    if (i & dj) continue;
                    u32 const e = i * 32;                          // lane bits do not matter for k >= 64
                    bool const desc = (k == 256) || ((e & k) != 0) ;
                    u32 const lo = min(key[i], key[i | dj]), hi = max(key[i], key[i | dj]);
                    key[i] = desc ? hi : lo; key[i | dj] = desc ? lo : hi;
                }
            } else {
                #pragma unroll
                for (u32 i = 0; i < 8; i++) {
                    u32 const e = i * 32 + lane;
                    u32 const other = __shfl_xor_sync(0xFFFFFFFFu, key[i], j);
                    bool const asc = (k == 256) ? false : ((e & k) == 0);
                    bool const takeMin = ((e & j) == 0) == asc;
                    key[i] = takeMin ? min(key[i], other) : max(key[i], other);
                }
            }
        }
    }
}

// Warp-cooperative variant (one warp builds one table; all operands in shared memory).
// store: HNode[2*256+2]; lenOf: u8[256]; firstVal: u32[16+16] (second half: running per-length counters).
// Returns max code length or an error (warp-uniform).
__device__ inline u64 warp_huf_build_ctable(u32* ctable, const u32* count, u32 msv, u32 maxBits,
                                            HNode* store, u8* lenOf, u32* firstVal)
{
    unsigned const lane = threadIdx.x & 31u;
    HNode* const nd = store + 1;
    if (!maxBits) maxBits = HUF_DEF_TLOG;
    if (msv > HUF_MAX_SV) return err(E_MSV_TOO_LARGE);
    {   // HUF_sort (huf_compress.c:307-329): decreasing count, ties by increasing symbol == decreasing (count << 8 | 255 - symbol)
        u32 key[8];
        #pragma unroll
        for (u32 i = 0; i < 8; i++) { u32 const s = i * 32 + lane; key[i] = ((s <= msv ? count[s] : 0u) << 8) | (255u - s); }
        warp_sort256_desc(key, lane);
        u64* const z = reinterpret_cast<u64*>(store);            // HNode is 8 bytes: {count, parent, sym, len}
        #pragma unroll
        for (u32 i = 0; i < 8; i++) {
            HNode h; h.count = key[i] >> 8; h.parent = 0; h.sym = (u8)(255u - (key[i] & 0xFFu)); h.len = 0;
            nd[i * 32 + lane] = h;
        }
        if (lane == 0) z[0] = 0;
        for (u32 i = 257 + lane; i < 2 * 256 + 2; i += 32) z[i] = 0;
    }
    __syncwarp();
    u32 mb = 0; int last = 0;
    if (lane == 0) {
        last = (int)msv; while (nd[last].count == 0) last--;
        int fresh = 256, leaf = last, root = fresh + leaf - 1, inner = fresh;
        nd[fresh].count = nd[leaf].count + nd[leaf - 1].count;
        nd[leaf].parent = nd[leaf - 1].parent = (u16)fresh;
        fresh++; leaf -= 2;
        for (int n = fresh; n <= root; n++) nd[n].count = 1u << 30;
        nd[-1].count = 1u << 31;
        u32 cl = nd[leaf].count, ci = nd[inner].count;           // heads of the two queues, kept in registers
        while (fresh <= root) {
            int a, b2; u32 ca, cb;
            if (cl < ci) { a = leaf--; ca = cl; cl = nd[leaf].count; } else { a = inner++; ca = ci; ci = nd[inner].count; }
            if (cl < ci) { b2 = leaf--; cb = cl; cl = nd[leaf].count; } else { b2 = inner++; cb = ci; ci = nd[inner].count; }
            nd[fresh].count = ca + cb;
            if (inner == fresh) ci = ca + cb;                    // the new node is the head of the inner queue
            nd[a].parent = nd[b2].parent = (u16)fresh;
            fresh++;
        }
        nd[root].len = 0;
        for (int n = root - 1; n >= 256; n--) nd[n].len = (u8)(nd[nd[n].parent].len + 1);
    }
    last = __shfl_sync(0xFFFFFFFFu, last, 0);
    __syncwarp();
    for (int n = (int)lane; n <= last; n += 32) nd[n].len = (u8)(nd[nd[n].parent].len + 1);
    __syncwarp();
    if (lane == 0) {
        mb = d_huf_limit_depth(nd, (u32)last, maxBits);
        if (mb <= HUF_MAX_TLOG) {
            u32 perLen[HUF_MAX_TLOG + 1];
            for (u32 i = 0; i <= HUF_MAX_TLOG; i++) perLen[i] = 0;
            for (int n = 0; n <= last; n++) perLen[nd[n].len]++;
            u32 v = 0;
            for (u32 i = 0; i <= HUF_MAX_TLOG; i++) firstVal[i] = 0;
            for (int n = (int)mb; n > 0; n--) { firstVal[n] = v; v = (v + perLen[n]) >> 1; }
        }
    }
    mb = __shfl_sync(0xFFFFFFFFu, mb, 0);
    if (mb > HUF_MAX_TLOG) return err(E_GENERIC);
    if (lane < 16) firstVal[16 + lane] = 0;
    __syncwarp();
    for (u32 s = lane; s <= msv; s += 32) lenOf[nd[s].sym] = nd[s].len;
    __syncwarp();
    // value = first value of the length + number of earlier symbols of that length (huf_compress.c:401-407), 32 symbols a round
    u32* const seen = firstVal + 16;
    for (u32 s0 = 0; s0 < 256; s0 += 32) {
        u32 const s = s0 + lane;
        u32 const len = s <= msv ? lenOf[s] : 31u;
        u32 const peers = __match_any_sync(0xFFFFFFFFu, len);
        u32 val = 0;
        if (s <= msv) val = ((firstVal[len] + seen[len] + __popc(peers & ((1u << lane) - 1))) & 0xFFFF) | (len << 16);
        __syncwarp();
        if (s <= msv && (peers >> lane) == 1u) seen[len] += __popc(peers);
        ctable[s] = val;
        __syncwarp();
    }
    return mb;
}

// HUF_compressWeights, one lane.  wksp: >= 160 u32 of scratch.
__device__ inline u64 d_huf_compress_weights(u8* out, u64 cap, const u8* w, u64 n, u32* wksp)
{
    unsigned* const count = wksp;                     // 13
    short* const norm = (short*)(wksp + 16);          // 13 shorts
    u32* const ct = wksp + 24;                        // 1 + 32 + 2*13 = 59
    u16* const cellSym = (u16*)(wksp + 84);           // 64 u16
    u32* const start = wksp + 116;                    // 15
    unsigned msv = HUF_MAX_TLOG, tl = 6, best = 0;
    if (n <= 1) return 0;
    for (unsigned i = 0; i <= HUF_MAX_TLOG; i++) count[i] = 0;
    for (u64 i = 0; i < n; i++) count[w[i]]++;
    while (!count[msv]) msv--;
    for (unsigned s = 0; s <= msv; s++) best = count[s] > best ? count[s] : best;
    if (best == n) return 1;
    if (best == 1) return 0;
    tl = d_optimal_tablelog(tl, n, msv, 2);
    u64 r = d_normalize(norm, tl, count, n, msv); if (is_err(r)) return r;
    r = d_write_ncount(out, cap, norm, msv, tl); if (is_err(r)) return r;
    u64 const o = r;
    d_build_ctable_serial(ct, norm, msv, tl, cellSym, start);
    r = d_fse_encode_serial(out + o, cap - o, w, n, ct);
    if (r == 0) return 0;
    return o + r;
}

// HUF_writeCTable, one lane.  `out` has room for min(cap,136) bytes where cap is the caller's capacity.
__device__ inline u64 d_huf_write_ctable(u8* out, u64 cap, const u32* ctable, unsigned msv, unsigned huffLog, u32* wksp)
{
    u8* const weight = (u8*)(wksp + 160);             // 256 bytes
    if (msv > HUF_MAX_SV) return err(E_MSV_TOO_LARGE);
    for (unsigned n = 0; n < msv; n++) {
        unsigned const len = (ctable[n] >> 16) & 0xFF;
        weight[n] = (u8)(len ? huffLog + 1 - len : 0);
    }
    {   u64 const h = d_huf_compress_weights(out + 1, cap - 1, weight, msv, wksp);
        if (is_err(h)) return h;
        if ((h > 1) & (h < msv / 2)) { out[0] = (u8)h; return h + 1; }
    }
    if (msv > 128) return err(E_GENERIC);
    if (((msv + 1) / 2) + 1 > cap) return err(E_DST_TOO_SMALL);
    out[0] = (u8)(128 + (msv - 1));
    weight[msv] = 0;
    for (unsigned n = 0; n < msv; n += 2) out[(n / 2) + 1] = (u8)((weight[n] << 4) + weight[n + 1]);
    return ((msv + 1) / 2) + 1;
}

// ---- warp-cooperative HUF_writeCTable (huf_compress.c:114-147) with HUF_compressWeights (:63-103) ----
// Same results as the one-lane versions above.  The weight histogram, the two tANS chains (lane 0: even indices /
// state 1, lane 1: odd indices / state 2 -- fse_compress.c:572-606) and the bit packing (prefix scan of the emitted
// (value, nbBits) records) use the warp; the 13-symbol normalisation, NCount and CTable stay on lane 0.
// out: shared memory, 4-byte aligned, at least `cap` (<= 136) bytes.  wksp: >= 360 u32.  All lanes call; result uniform.
__device__ inline u64 warp_huf_compress_weights(u8* out, u64 cap, const u8* w, u32 n, u32* wksp)
{
    unsigned const lane = threadIdx.x & 31u;
    unsigned* const count = wksp;                     // 13
    short* const norm = (short*)(wksp + 16);          // 13 shorts
    u32* const ct = wksp + 24;                        // 1 + 32 + 2*13 = 59
    u16* const cellSym = (u16*)(wksp + 84);           // 64 u16
    u32* const start = wksp + 116;                    // 15
    u16* const rec = (u16*)(wksp + 224);              // 256 records (value | nbBits << 8)
    if (n <= 1) return 0;
    if (lane < 16) count[lane] = 0;
    __syncwarp();
    for (u32 i = lane; i < n; i += 32) atomicAdd(&count[w[i]], 1u);
    __syncwarp();
    unsigned msv = HUF_MAX_TLOG, best = 0;
    while (!count[msv]) msv--;
    for (unsigned s = 0; s <= msv; s++) best = count[s] > best ? count[s] : best;
    if (best == n) return 1;
    if (best == 1) return 0;
    unsigned const tl = d_optimal_tablelog(6, n, msv, 2);
    u64 r = 0;
    if (lane == 0) {
        r = d_normalize(norm, tl, count, n, msv);
        if (!is_err(r)) r = d_write_ncount(out, cap, norm, msv, tl);
        if (!is_err(r)) d_build_ctable_serial(ct, norm, msv, tl, cellSym, start);
    }
    r = __shfl_sync(0xFFFFFFFFu, r, 0);
    if (is_err(r)) return r;
    u32 const o = (u32)r;
    __syncwarp();
    if (n <= 2) return 0;                             // fse_compress.c:563
    u64 const scap = cap - o;
    if (scap <= 8) return 0;                          // BIT_initCStream refuses (bitstream.h:190)
    CtView const c = ct_view(ct);
    u32 state = 0;
    if (lane < 2) {                                   // the two interleaved chains
        bool seeded = false;
        int const first = (int)(((n - 1) & 1u) == lane ? n - 1 : n - 2);
        for (int i = first; i >= 0; i -= 2) {
            u32 const sym = w[i];
            if (!seeded) { state = enc_seed(c, sym); seeded = true; rec[i] = 0; }
            else {
                u32 const nb = (state + c.tt[2 * sym + 1]) >> 16;
                rec[i] = (u16)((state & ((1u << nb) - 1)) | (nb << 8));
                state = c.next[(state >> nb) + c.tt[2 * sym]];
            }
        }
    }
    u32 const s0 = __shfl_sync(0xFFFFFFFFu, state, 0), s1 = __shfl_sync(0xFFFFFFFFu, state, 1);
    // zero the stream area, then OR every record at its bit offset (aligned words of the shared buffer)
    for (u32 i = o + lane; i < (u32)cap; i += 32) out[i] = 0;
    __syncwarp();
    u64 const a0 = reinterpret_cast<u64>(out + o);
    u32* const words = reinterpret_cast<u32*>(a0 & ~3ull);
    u32 const limitWords = (u32)(((a0 & 3) + scap + 3) / 4);
    u32 running = 8 * (u32)(a0 & 3);
    u32 const bit0 = running;
    u32 const R = n + 3;
    for (u32 k0 = 0; k0 < R; k0 += 32) {
        u32 const k = k0 + lane;
        u32 e = 0;
        if (k < n) e = rec[n - 1 - k];
        else if (k == n) e = (s1 & ((1u << tl) - 1)) | (tl << 8);        // flush state 2, then state 1, then the end mark
        else if (k == n + 1) e = (s0 & ((1u << tl) - 1)) | (tl << 8);
        else if (k == n + 2) e = 1u | (1u << 8);
        u32 const nb = e >> 8, val = e & 0xFF;
        u32 incl = nb;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 const t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (unsigned)d) incl += t; }
        u32 const sum = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (nb) {
            u32 const off = running + incl - nb;
            u32 const wi = off >> 5, sh = off & 31;
            if (wi < limitWords) atomicOr(&words[wi], val << sh);
            if (sh + nb > 32 && wi + 1 < limitWords) atomicOr(&words[wi + 1], val >> (32 - sh));
        }
        running += sum;
    }
    __syncwarp();
    u32 const totalBits = running - bit0;
    if ((u64)(totalBits >> 3) >= scap - 8) return 0;  // bitstream.h:246,258
    return (u64)o + ((totalBits + 7) >> 3);
}

__device__ inline u64 warp_huf_write_ctable(u8* out, u64 cap, const u32* ctable, unsigned msv, unsigned huffLog, u32* wksp)
{
    unsigned const lane = threadIdx.x & 31u;
    u8* const weight = (u8*)(wksp + 160);             // 256 bytes
    if (msv > HUF_MAX_SV) return err(E_MSV_TOO_LARGE);
    for (unsigned n = lane; n < msv; n += 32) {
        unsigned const len = (ctable[n] >> 16) & 0xFF;
        weight[n] = (u8)(len ? huffLog + 1 - len : 0);
    }
    __syncwarp();
    {   u64 const h = warp_huf_compress_weights(out + 1, cap - 1, weight, msv, wksp);
        if (is_err(h)) return h;
        if ((h > 1) & (h < msv / 2)) { if (lane == 0) out[0] = (u8)h; __syncwarp(); return h + 1; }
    }
    if (msv > 128) return err(E_GENERIC);
    if (((msv + 1) / 2) + 1 > cap) return err(E_DST_TOO_SMALL);
    __syncwarp();
    if (lane == 0) { out[0] = (u8)(128 + (msv - 1)); weight[msv] = 0; }
    __syncwarp();
    for (unsigned n = 2 * lane; n < msv; n += 64) out[(n / 2) + 1] = (u8)((weight[n] << 4) + weight[n + 1]);
    __syncwarp();
    return ((msv + 1) / 2) + 1;
}

}  // namespace fseb
