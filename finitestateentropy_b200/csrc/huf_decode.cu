// huf_decode.cu -- batched Huff0 4-stream decode for sm_100a (HBM-bound integer path, no tensor cores).
//
// Replaces, for a whole batch of independent blocks, the CPU chain
//   HUF_decompress            lib/huf_decompress.c:1056-1081  (raw / RLE / decoder choice)
//   HUF_readStats             lib/entropy_common.c:154-215
//   HUF_readDTableX1          lib/huf_decompress.c:118-185
//   HUF_decompress4X1_usingDTable_internal_body  lib/huf_decompress.c:262-354
// Decoded bytes and error verdicts are those of the single-symbol (X1) decoder; X1 vs X2 is a CPU
// speed heuristic only (lib/huf_decompress.c:1029-1051) and both regenerate identical bytes.
//
// B200 mapping ("lane per stream, table per bank"):
//   * one CTA = up to 64 consecutive blocks = 256 streams = 256 threads; a warp holds the SAME stream
//     index of 32 DIFFERENT blocks, so the 32 lanes of a table look-up hit 32 different tables;
//   * per-block decode tables are interleaved by column: u16 main[512][64] is indexed by the top 9
//     bits of the bit window, u16 sub[192][64] by the top tableLog bits for the few windows that start
//     a code longer than 9 bits (Huff0 numbers the longest codes from 0, so those windows are exactly
//     the values below a per-block threshold T).  Column j of a row sits in bank j/2, and a warp holds
//     the even (or the odd) columns, so every look-up is bank-conflict free whatever the 32 indices;
//   * the choice main/sub is a select, not a branch: the hot loop has no divergent control flow;
//   * each lane keeps a 64-bit left-aligned bit window in two registers; every two symbols a
//     predicated block merges the next 32-bit word, taken from a private 16-word shared-memory ring the
//     lane fills itself with cp.async (global -> shared, no register dependency, waited one visit later);
//   * output: 16 symbols are packed into one 16-byte store per lane into its quarter of the block.  Global traffic
//     is inherently per-lane (every lane owns a different stream), so the number of L1 wavefronts -- one per lane per
//     request -- is what bounds this kernel: 16-byte requests both ways keep it at 1 wavefront per 16 bytes;
//   * blocks whose long-code region exceeds 192 windows ("hard", e.g. near-flat 256-symbol alphabets),
//     unaligned segments and ragged tails take a slower per-symbol loop with a canonical-code search;
//   * the grid is shaped so that every SM gets the same number of blocks per round (a round = 2 CTAs
//     per SM): blocksPerCta is chosen <= 64 from the batch size, see launch_huf_decode().
#include "common.cuh"
#include "huf_dev.cuh"

namespace fseb {
namespace hufd {

constexpr int G = 64;             // block columns per CTA
constexpr int THREADS = 4 * G;    // one lane per stream
constexpr int MAIN_BITS = 9;
constexpr int MAIN_ROWS = 1 << MAIN_BITS;
constexpr int SUB_ROWS = 192;
constexpr int RING = 16;          // 32-bit words of stream look-ahead per lane
constexpr u32 NOERR = 0xFFFFFFFFu;
constexpr unsigned FULL = 0xFFFFFFFFu;

struct __align__(16) Smem {
    u16 main[MAIN_ROWS][G];       // 64 KB   first-level table, column = block
    u16 sub[SUB_ROWS][G];         // 24 KB   long-code windows (index < T); hard blocks park their sorted symbol list here
    uint4 ring[RING / 4][THREADS]; // 16 KB   per-lane stream chunks (16 bytes = 4 words, highest address first when consumed); table-build scratch before the streams start
    u16 rankEnd[HUF_MAX_TLOG + 2][G];   // end (exclusive) of weight w's range in tableLog-bit index space
    u16 listStart[HUF_MAX_TLOG + 2][G]; // first position of weight w in the sorted symbol list
    u16 longT[G];                 // number of tableLog-bit windows that start a code longer than 9 bits
    u8  tlog[G];
    u8  kind[G];                  // 0 = Huffman, 1 = raw copy, 2 = RLE, 3 = done/skip
    u8  hard[G];
    u32 status[G];                // NOERR or (stage<<8 | error code), smallest wins
    u32 hsize[G];                 // header bytes
    u32 rankStats[THREADS / 32][HUF_MAX_TLOG + 1];
    u16 rankRun[THREADS / 32][HUF_MAX_TLOG + 2];
};

// canonical-code look-up in tableLog-bit index space -> (nbBits | symbol << 8); `list(k)` yields the k-th symbol in code order
template <typename ListFn>
__device__ __forceinline__ u32 canon_lookup(const Smem& sm, int blk, u32 idx, u32 tl, u32& w, ListFn list)
{
    while (idx >= sm.rankEnd[w][blk]) w++;
    u32 const first = (w == 1) ? 0u : sm.rankEnd[w - 1][blk];
    u32 const k = sm.listStart[w][blk] + ((idx - first) >> (w - 1));
    return (tl + 1 - w) | (list(k) << 8);
}

// Builds the tables of block column `blk` with one warp.
__device__ void setup_block(Smem& sm, int blk, const u8* csrc, u64 csize, int warp)
{
    unsigned const lane = lane_id();
    u8* const weights = reinterpret_cast<u8*>(&sm.ring[0][0]) + warp * 512;   // per-warp scratch inside the (not yet used) ring
    u8* const sorted = weights + 256;                                          // symbols ordered by (weight asc, symbol asc) == code order
    u32 nbSym = 0, tl = 0; u64 h = 0;
    if (lane == 0) {
        h = d_huf_read_stats(weights, 256, sm.rankStats[warp], &nbSym, &tl, csrc, csize);
        if (!is_err(h) && tl > HUF_MAX_TLOG) h = err(E_TLOG_TOO_LARGE);      // huf_decompress.c:143
        if (!is_err(h) && h >= csize) h = err(E_SRC_WRONG);                   // huf_decompress.c:426
        if (!is_err(h)) {
            // rank ranges (huf_decompress.c:151-156): weight w covers 2^(w-1) cells per symbol, longest codes first
            u32 acc = 0, pos = 0;
            sm.rankEnd[0][blk] = 0; sm.listStart[0][blk] = 0;
            for (u32 w = 1; w <= tl; w++) {
                sm.listStart[w][blk] = (u16)pos; sm.rankRun[warp][w] = (u16)pos;
                pos += sm.rankStats[warp][w];
                acc += sm.rankStats[warp][w] << (w - 1);
                sm.rankEnd[w][blk] = (u16)(acc > 0xFFFF ? 0xFFFF : acc);
            }
            sm.rankEnd[tl][blk] = (u16)(1u << tl);
            sm.rankEnd[tl + 1][blk] = 0xFFFF;
            u32 const T = (tl > MAIN_BITS) ? sm.rankEnd[tl - MAIN_BITS][blk] : 0u;     // weights 1..tl-9 <=> code length > 9
            sm.longT[blk] = (u16)T;
            sm.hard[blk] = (u8)(T > SUB_ROWS);
            sm.tlog[blk] = (u8)tl;
            sm.hsize[blk] = (u32)h;
        } else {
            atomicMin(&sm.status[blk], (u32)(0u << 8 | (u32)(0 - h)));
        }
    }
    h = __shfl_sync(FULL, h, 0);
    if (is_err(h)) return;
    nbSym = __shfl_sync(FULL, nbSym, 0);
    tl = __shfl_sync(FULL, tl, 0);
    __syncwarp();
    // sorted symbol list: stable by weight, then symbol order (huf_decompress.c:158-183 fills cells in that order)
    for (u32 base = 0; base < nbSym; base += 32) {
        u32 const s = base + lane;
        u32 const w = (s < nbSym) ? weights[s] : 0u;
        u32 const peers = __match_any_sync(FULL, w);
        if (w) sorted[sm.rankRun[warp][w] + __popc(peers & ((1u << lane) - 1))] = (u8)s;
        __syncwarp();
        if (w && (peers >> lane) == 1u) sm.rankRun[warp][w] = (u16)(sm.rankRun[warp][w] + __popc(peers));   // highest lane of the group
        __syncwarp();
    }
    auto list = [&](u32 k) -> u32 { return sorted[k]; };
    // first-level table: cell = nbBits | symbol<<8 (cells that start a longer code are never read: the select picks `sub`)
    {   u32 w = 1;
        for (u32 i = 0; i < MAIN_ROWS / 32; i++) {
            u32 const idx9 = lane * (MAIN_ROWS / 32) + i;
            u32 const idx = (tl >= MAIN_BITS) ? (idx9 << (tl - MAIN_BITS)) : (idx9 >> (MAIN_BITS - tl));
            u32 const e = canon_lookup(sm, blk, idx, tl, w, list);
            sm.main[idx9][blk] = (u16)((e & 0xFF) <= MAIN_BITS ? e : 0u);
        }
    }
    u32 const T = sm.longT[blk];
    if (T <= SUB_ROWS) {
        for (u32 idx = lane; idx < T; idx += 32) { u32 w = 1; sm.sub[idx][blk] = (u16)canon_lookup(sm, blk, idx, tl, w, list); }
    } else {
        for (u32 r = lane; r < 128; r += 32) sm.sub[r][blk] = (u16)(sorted[2 * r] | (sorted[2 * r + 1] << 8));   // hard block: park the code-ordered symbol list, 2 per cell
    }
    __syncwarp();
}

__device__ __forceinline__ u32 lds_u16(u32 addr) { u16 v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr)); return v; }

__global__ void __launch_bounds__(THREADS, 2)
huf_decode_kernel(BatchGeom g, u8* __restrict__ dst, const u8* __restrict__ cbuf, const u64* __restrict__ csizes,
                  u64* __restrict__ results, const u8* __restrict__ orig, u32 flags, u32 gEff)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    int const tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    u32 const blk0 = blockIdx.x * gEff;
    u32 const blkEnd = min(blk0 + gEff, g.nBlocks);                         // this CTA owns blocks [blk0, blkEnd)

    if (tid < G) { sm.status[tid] = NOERR; sm.kind[tid] = 3; sm.hsize[tid] = 0; sm.tlog[tid] = 0; sm.longT[tid] = 0; sm.hard[tid] = 0; }
    __syncthreads();

    // ---- classify blocks and build tables: warp w handles columns w, w+8, ... ----
    for (int j = warp; j < G; j += THREADS / 32) {
        u32 const b = blk0 + j;
        if (b >= blkEnd) continue;                                      // warp-uniform
        u64 const n = block_len(g, b);
        u64 const cs = csizes[b];
        int kind;
        if (flags & 1u) kind = 0;                                       // HUF_decompress4X1 semantics: always a Huffman block
        else if (is_err(cs)) { kind = 3; if (lane == 0) results[b] = cs; }   // propagated compressor error
        else if (cs == 0) { kind = orig ? 1 : 3; if (lane == 0) results[b] = orig ? n : 0; }   // stored raw by the harness (bench.c:393-397)
        else if (n == 0) { kind = 3; if (lane == 0) results[b] = err(E_DST_TOO_SMALL); }        // huf_decompress.c:1063
        else if (cs > n) { kind = 3; if (lane == 0) results[b] = err(E_CORRUPT); }              // :1064
        else if (cs == n) kind = 1;                                                                // :1065
        else if (cs == 1) kind = 2;                                                                // :1066
        else kind = 0;
        if (lane == 0) sm.kind[j] = (u8)kind;
        if (kind == 0) setup_block(sm, j, cbuf + (u64)b * g.slot, cs, warp);
    }
    __syncthreads();

    // ---- per-lane stream set-up: thread -> (block column, stream) ----
    int const col = 2 * lane + (warp >> 2);          // a warp sees the 32 even (or odd) columns = 32 distinct banks
    int const strm = warp & 3;
    u32 const b = blk0 + col;
    bool const live = (b < blkEnd) && sm.kind[col] == 0 && sm.status[col] == NOERR;
    u32 const n = live ? block_len(g, b) : 0;
    u32 const seg = (n + 3) / 4;
    u32 segLen = 0;                                  // symbols this lane must produce
    u8* outp = dst + (u64)b * g.blockSize + (u64)strm * seg;
    u32 hi = 0, lo = 0, nw = 0, r = 0, k = 0, q = 0;
    u64 chunkTop = 0, sBegin = 0;                    // address just above chunk 0 ; first byte of the stream
    u32 const tl = sm.tlog[col];

    u32 fullChunks = 0;                              // chunks 0..fullChunks-1 lie entirely inside the stream (plain 16-byte copies)
    u32 const ringLo = (u32)__cvta_generic_to_shared(&sm.ring[0][tid]);   // this thread's 16 bytes of chunk slot 0
    // Chunk qq = the aligned 16 bytes ending at chunkTop - 16*qq; it holds stream words 4qq..4qq+3 in descending
    // address order; bytes below the stream start read as 0 (the reference's reader pads the same way).
    auto stage_sync = [&](u32 qq) {                  // boundary / out-of-stream chunks and the initial fill: through registers
        u64 const top = chunkTop - 16ull * qq;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (top > sBegin && 16ull * qq < chunkTop) {
            v = __ldg(reinterpret_cast<const uint4*>(top - 16));
            if (top - 16 < sBegin) {
                u32 const z = (u32)(sBegin - (top - 16));        // 1..15 leading bytes to clear
                u32* p = reinterpret_cast<u32*>(&v);
                #pragma unroll
                for (int i = 0; i < 4; i++) {
                    int const zb = (int)z - 4 * i;
                    if (zb >= 4) p[i] = 0; else if (zb > 0) p[i] &= 0xFFFFFFFFu << (8 * zb);
                }
            }
        }
        sm.ring[qq & (RING / 4 - 1)][tid] = v;
    };
    auto stage_async = [&](u32 qq) {                 // steady state: ONE 16-byte cp.async per chunk (one L1 wavefront), no register dependency
        u64 const top = chunkTop - 16ull * qq;
        u32 const d0 = ringLo + (qq & (RING / 4 - 1)) * (THREADS * 16);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.commit_group;" :: "r"(d0), "l"(top - 16) : "memory");
    };
    auto ring_word = [&](u32 j) -> u32 {             // stream word j (descending addresses): chunk j/4, word 3 - j%4
        return reinterpret_cast<const u32*>(&sm.ring[(j >> 2) & (RING / 4 - 1)][tid])[3 - (j & 3)];
    };

    if (live) {
        const u8* const cs0 = cbuf + (u64)b * g.slot;
        u64 const cs = csizes[b];
        const u8* const pay = cs0 + sm.hsize[col];
        u64 const psize = cs - sm.hsize[col];
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
                        sBegin = (u64)(pay + off);
                        u64 const e = sBegin + len;                                  // one past the last byte
                        chunkTop = ((e - 1) & ~15ull) + 16;
                        segLen = (strm < 3) ? seg : n - 3 * seg;
                    }
                }
            }
        }
        if (code) atomicMin(&sm.status[col], (u32)((1u + strm) << 8 | code));
    }
    __syncthreads();                                  // init verdicts of all four streams are in; table-build scratch (ring) is free now
    bool const go = live && sm.status[col] == NOERR;
    if (!go) segLen = 0;
    // window initialisation (needs the ring, hence after the barrier)
    if (go) {
        const u8* const cs0 = cbuf + (u64)b * g.slot;
        u64 const cs = csizes[b];
        const u8* const pay = cs0 + sm.hsize[col];
        u64 const psize = cs - sm.hsize[col];
        u32 const l1 = rd16(pay), l2 = rd16(pay + 2), l3 = rd16(pay + 4);
        u32 const l4 = (u32)(psize - 6 - l1 - l2 - l3);
        u32 const off = 6 + (strm > 0 ? l1 : 0) + (strm > 1 ? l2 : 0) + (strm > 2 ? l3 : 0);
        u32 const len = strm == 0 ? l1 : strm == 1 ? l2 : strm == 2 ? l3 : l4;
        u8 const last = pay[off + len - 1];
        u64 const e = sBegin + len;
        u32 const c0 = (u32)(8 * (chunkTop - e)) + (8 - hibit(last));      // garbage above the stream + zero padding + end mark
        fullChunks = (u32)((chunkTop - ((sBegin + 15) & ~15ull)) >> 4);
        stage_sync(0); stage_sync(1); stage_sync(2); stage_sync(3);
        q = 4;
        k = c0 >> 5; r = c0 & 31;
        u32 const w0 = ring_word(k), w1 = ring_word(k + 1);
        nw = ring_word(k + 2);
        hi = __funnelshift_l(w1, w0, r); lo = w1 << r;
    }

    // ---- decode ----
    u32 const sMain = (u32)__cvta_generic_to_shared(&sm.main[0][col]);     // + row * 128
    u32 const sSub = (u32)__cvta_generic_to_shared(&sm.sub[0][col]);
    u32 const ringBase = (u32)__cvta_generic_to_shared(&sm.ring[0][0]);
    // ring cursor: byte offset (chunkSlot * 4096 + tid * 16 + word * 4) of the word `nw` was read from; a refill moves it
    // first: down by 4 inside a chunk, or to word 3 of the next chunk slot.
    // The word index k = floor(consumed / 32) is not carried through the hot loop: it is recovered from the cursor
    // and the staging counter q (k + 2 is in [4q-16, 4q-1] because a chunk is staged only when 4q <= k + 14).
    u32 roff = (((k + 2) >> 2) & (RING / 4 - 1)) * (THREADS * 16) + tid * 16 + (3 - ((k + 2) & 3)) * 4;
    u32 const T = sm.longT[col];
    u32 const shTl = 32 - (tl ? tl : 1);
    u32 const Thi = T << shTl;                        // window values below this start a code longer than 9 bits (T <= 192 on the fast path)
    u32 const mulTl = 1u << (tl ? tl : 1);            // (hi * 2^tl) >> 32 == hi >> (32 - tl), computed on the FMA pipe
    bool const hardBlk = sm.hard[col] != 0;
    auto word_index = [&]() -> u32 {                  // k from the ring cursor
        u32 const slot = (roff >> 12) * 4 + (3 - ((roff >> 2) & 3));     // == (k + 2) mod 16
        u32 const lo16 = 4 * q - 16;                  // smallest possible k + 2
        return lo16 + ((slot - lo16) & (RING - 1)) - 2;
    };

    // Keep the ring ahead of the consumer.  Called every 16 symbols (<= 6 words consumed in between).  Normal case: one
    // chunk (4 words) is staged asynchronously when its slots are free (4q <= k+14).  If that leaves less than the
    // worst case of the next interval in the ring (4q - k < 13; only for data averaging > 8 bits/symbol), more chunks
    // are staged through registers and every pending copy is awaited.  Otherwise the words needed before the next call
    // (<= k+8 < 4q-4) are older than the newest copy group of this warp -- cp.async groups are tracked per WARP
    // (LDGDEPBAR / DEPBAR.LE), at most one group per call -- so waiting for all but the newest is enough.
    // (Committing two groups per call and waiting for "all but the newest" exposed a DRAM latency per call.)
    u64 gsrc = chunkTop - 16ull * (q + 1);            // global address of chunk q (valid while q < fullChunks)
    u32 sdst = (q & (RING / 4 - 1)) * (THREADS * 16);  // its ring slot (byte offset from ringLo)
    auto top_up = [&]() {
        u32 const kk = word_index();
        u32 v = 4 * q - kk;
        if (v <= 14) {
            if (q < fullChunks) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.commit_group;" :: "r"(ringLo + sdst), "l"(gsrc) : "memory");
            else stage_sync(q);
            q++; gsrc -= 16; sdst = (sdst + THREADS * 16) & (RING * THREADS * 4 - 1); v += 4;
        }
        if (__builtin_expect(v < 13, 0)) {            // emergency: the consumer outruns one chunk per call
            while (4 * q - kk <= 14) { stage_sync(q); q++; gsrc -= 16; sdst = (sdst + THREADS * 16) & (RING * THREADS * 4 - 1); }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else if (q >= fullChunks) asm volatile("cp.async.wait_group 0;" ::: "memory");   // tail of the stream: chunks come through registers
        else asm volatile("cp.async.wait_group 1;" ::: "memory");
    };
    // One symbol, branch-free: E = nbBits | symbol << 8, window advanced.  Integer work is split between the ALU pipe
    // (and / setp / funnel shifts) and the FMA pipe (mad.hi / mad.lo do the two shifts-and-adds that form the addresses):
    // both pipes issue one warp instruction every two cycles, and the ALU pipe is what bounds this loop.
#define HUFD_LOOKUP(E) do { \
        u32 aM_, aS_, i12_, tM_; u16 e16_; \
        asm volatile("{\n\t.reg .pred p;\n\t" \
            "and.b32 %3, %5, 0xFF800000;\n\t"            /* top 9 bits */ \
            "mad.hi.u32 %1, %3, 65536, %6;\n\t"           /* (idx9 << 7) + sMain */ \
            "setp.lt.u32 p, %5, %8;\n\t" \
            "mul.hi.u32 %4, %5, %9;\n\t"                   /* top tableLog bits */ \
            "@p mad.lo.u32 %1, %4, 128, %7;\n\t"          /* (idx << 7) + sSub */ \
            "ld.shared.u16 %0, [%1];\n\t}" \
            : "=h"(e16_), "=r"(aM_), "=r"(aS_), "=r"(tM_), "=r"(i12_) \
            : "r"(hi), "r"(sMain), "r"(sSub), "r"(Thi), "r"(mulTl)); \
        E = e16_; \
        hi = __funnelshift_l(lo, hi, E); lo = __funnelshift_l(0, lo, E); \
    } while (0)
    // merge the next word when the pair crossed a 32-bit boundary (bit 5 of the running count); all predicated
#define HUFD_REFILL() asm volatile("{\n\t" \
        ".reg .pred p, w0;\n\t.reg .b32 t, u, v;\n\t" \
        "and.b32 t, %4, 32;\n\t" \
        "setp.ne.u32 p, t, 0;\n\t" \
        "and.b32 %4, %4, 31;\n\t" \
        "shf.l.wrap.b32 t, %2, 0, %4;\n\t" \
        "@p or.b32 %0, %0, t;\n\t" \
        "@p shl.b32 %1, %2, %4;\n\t" \
        "and.b32 u, %3, 12;\n\t" \
        "setp.eq.u32 w0, u, 0;\n\t"                      /* last word of the chunk consumed: jump to word 3 of the next slot */ \
        "add.u32 v, %3, -4;\n\t" \
        "mad.lo.u32 u, %5, 1, %3;\n\t" \
        "@w0 and.b32 v, u, %6;\n\t" \
        "@p mov.b32 %3, v;\n\t" \
        "@p add.u32 u, v, %7;\n\t" \
        "@p ld.shared.u32 %2, [u];\n\t" \
        "}" : "+r"(hi), "+r"(lo), "+r"(nw), "+r"(roff), "+r"(r) : "r"(THREADS * 16 + 12), "n"(RING * THREADS * 4 - 1), "r"(ringBase) : "memory")

    u32 pos = 0;
    bool const fastOk = go && !hardBlk && ((reinterpret_cast<u64>(outp) & 15) == 0);
    if (fastOk) {
        u32 const nIter = segLen >> 4;
        for (u32 it = 0; it < nIter; it++) {             // 16 symbols -> one 16-byte store per lane
            u32 o[4], e0, e1;
            #pragma unroll
            for (int h = 0; h < 4; h++) {
                if (h == 0) top_up();                    // every 16 symbols
                u32 t;
                HUFD_LOOKUP(e0); HUFD_LOOKUP(e1); r += e0 + e1; HUFD_REFILL();
                t = __byte_perm(e0, e1, 0x0051);                        // {sym0, sym1, x, x}
                HUFD_LOOKUP(e0); HUFD_LOOKUP(e1); r += e0 + e1; HUFD_REFILL();
                o[h] = __byte_perm(t, __byte_perm(e0, e1, 0x0051), 0x5410);
            }
            *reinterpret_cast<uint4*>(outp + pos) = make_uint4(o[0], o[1], o[2], o[3]);
            pos += 16;
        }
    }
    // ragged tails, unaligned segments and hard blocks: one symbol at a time
    {
        auto parked = [&](u32 kk) -> u32 { return (sm.sub[kk >> 1][col] >> (8 * (kk & 1))) & 0xFFu; };
        while (pos < segLen) {
            if ((pos & 7) == 0) top_up();
            u32 e;
            u32 const idx = hi >> shTl;
            if (idx < T) {
                if (hardBlk) { u32 w = 1; e = canon_lookup(sm, col, idx, tl, w, parked); }
                else e = sm.sub[idx][col];
            } else e = sm.main[hi >> (32 - MAIN_BITS)][col];
            hi = __funnelshift_l(lo, hi, e); lo = __funnelshift_l(0, lo, e);
            r += e & 0xFF;
            HUFD_REFILL();
            outp[pos++] = (u8)(e >> 8);
        }
    }
#undef HUFD_LOOKUP
#undef HUFD_REFILL

    asm volatile("cp.async.wait_all;" ::: "memory");
    // ---- verdict: every stream must be consumed exactly (huf_decompress.c:348-349) ----
    if (go) {
        u64 const consumed = 32ull * word_index() + (r & 31);
        u64 const expect = 8ull * (chunkTop - sBegin);
        if (consumed != expect) atomicMin(&sm.status[col], (u32)(5u << 8 | E_CORRUPT));
    }
    __syncthreads();
    if (tid < G) {
        u32 const bb = blk0 + tid;
        if (bb < blkEnd && sm.kind[tid] == 0) {
            u32 const st = sm.status[tid];
            results[bb] = (st == NOERR) ? (u64)block_len(g, bb) : err(st & 0xFF);
        }
    }
    // ---- raw / RLE blocks (huf_decompress.c:1065-1066; the harness' own 0-size convention, bench.c:393-402) ----
    for (int j = 0; j < G; j++) {
        int const kd = sm.kind[j];
        if (kd != 1 && kd != 2) continue;
        u32 const bb = blk0 + j;
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

cudaError_t launch_huf_decode(const BatchGeom& g, void* dst, const void* cbuf, const u64* csizes, u64* results,
                              const void* orig, cudaStream_t stream, u32 flags)
{
    static bool configured = false;
    static int numSMs = 148;
    size_t const smem = sizeof(hufd::Smem);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(hufd::huf_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    if (g.nBlocks == 0) return cudaSuccess;
    // balanced rounds: every lane decodes a whole stream, so a CTA's run time does not depend on how many
    // blocks it holds; give each SM the same number of blocks per round (2 CTAs resident per SM).
    u32 const slots = 2u * (u32)numSMs;
    u32 const rounds = (g.nBlocks + slots * hufd::G - 1) / (slots * hufd::G);
    u32 gEff = (g.nBlocks + slots * rounds - 1) / (slots * rounds);
    if (gEff > (u32)hufd::G) gEff = hufd::G;
    if (gEff < 1) gEff = 1;
    unsigned const grid = (g.nBlocks + gEff - 1) / gEff;
    hufd::huf_decode_kernel<<<grid, hufd::THREADS, smem, stream>>>(g, (u8*)dst, (const u8*)cbuf, csizes, results, (const u8*)orig, flags, gEff);
    return cudaGetLastError();
}

}  // namespace fseb
