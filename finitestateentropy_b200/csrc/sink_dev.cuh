// sink_dev.cuh -- single-lane forward bit writer and the 2-state tANS encoder on top of it.
//
// Wire format of BIT_CStream_t (lib/bitstream.h:57-63,183-260): bits are appended LSB-first and a
// single 1 closes the stream.  Capacity rule (:190-191,246,258): the stream is refused (size 0)
// when cap <= 8 or floor(totalBits/8) >= cap-8, whatever the flush schedule was.
// Encoder: FSE_compress_usingCTable_generic (lib/fse_compress.c:554-611, lib/fse.h:503-527).
// Used for the short serial jobs (Huffman weight headers, one-block API calls).
#pragma once
#include "common.cuh"

namespace fseb {

struct BitSink { u8* out; u64 cap; u64 acc; unsigned held; u64 nbytes; u64 nbits; bool usable; };

__device__ __forceinline__ void sink_open(BitSink& s, u8* dst, u64 cap)
{
    s.out = dst; s.cap = cap; s.acc = 0; s.held = 0; s.nbytes = 0; s.nbits = 0; s.usable = cap > 8;
}
__device__ __forceinline__ void sink_put(BitSink& s, u32 value, unsigned nb)
{
    if (nb == 0) return;
    s.acc |= ((u64)value & ((1ULL << nb) - 1)) << s.held;
    s.held += nb; s.nbits += nb;
    while (s.held >= 8) {
        if (s.nbytes < s.cap) s.out[s.nbytes] = (u8)s.acc;
        s.nbytes++; s.acc >>= 8; s.held -= 8;
    }
}
__device__ __forceinline__ u64 sink_close(BitSink& s)
{
    if (!s.usable) return 0;
    sink_put(s, 1, 1);
    if ((s.nbits >> 3) >= s.cap - 8) return 0;
    if (s.held) { s.out[s.nbytes] = (u8)s.acc; s.nbytes++; }
    return s.nbytes;
}

struct CtView { const u16* next; const u32* tt; unsigned tl; };
__device__ __forceinline__ CtView ct_view(const u32* ct)
{
    CtView v;
    v.tl = ((const u16*)ct)[0];
    v.next = ((const u16*)ct) + 2;
    v.tt = ct + 1 + (v.tl ? (1u << (v.tl - 1)) : 1);
    return v;
}
__device__ __forceinline__ u32 enc_seed(const CtView& c, u32 sym)
{
    u32 const dnb = c.tt[2 * sym + 1];
    u32 const nb = (dnb + (1u << 15)) >> 16;
    u32 const v = (nb << 16) - dnb;
    return c.next[(v >> nb) + c.tt[2 * sym]];
}
__device__ __forceinline__ u32 enc_step(BitSink& s, const CtView& c, u32 state, u32 sym)
{
    u32 const nb = (state + c.tt[2 * sym + 1]) >> 16;
    sink_put(s, state, nb);
    return c.next[(state >> nb) + c.tt[2 * sym]];
}

// byte symbols, two interleaved states (even index -> state 1, odd -> state 2)
__device__ inline u64 d_fse_encode_serial(u8* dst, u64 cap, const u8* in, u64 n, const u32* ct)
{
    CtView const c = ct_view(ct);
    BitSink s;
    u32 st[2] = { 0, 0 }; bool seeded[2] = { false, false };
    if (n <= 2) return 0;
    sink_open(s, dst, cap);
    if (!s.usable) return 0;
    for (u64 i = n; i-- > 0;) {
        unsigned const k = (unsigned)(i & 1);
        if (!seeded[k]) { st[k] = enc_seed(c, in[i]); seeded[k] = true; }
        else st[k] = enc_step(s, c, st[k], in[i]);
    }
    sink_put(s, st[1], c.tl);
    sink_put(s, st[0], c.tl);
    return sink_close(s);
}

}  // namespace fseb
