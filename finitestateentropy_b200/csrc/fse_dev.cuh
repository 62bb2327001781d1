// fse_dev.cuh -- device-side FSE statistics and table construction (one lane per table unless noted).
//
// Bit-exact restatement, for the GPU, of:
//   FSE_optimalTableLog      lib/fse_compress.c:316-342
//   FSE_normalizeCount (+M2) lib/fse_compress.c:348-494
//   FSE_writeNCount          lib/fse_compress.c:186-298
//   FSE_readNCount           lib/entropy_common.c:41-144
//   FSE_buildCTable          lib/fse_compress.c:66-169   (layout lib/fse.h:295,483-486)
//   FSE_buildDTable          lib/fse_decompress.c:71-126 (layout lib/fse.h:296,565-575; U16: lib/fseU16.c:78-82)
// These are O(alphabet)/O(table) integer routines with CPU-specific tie-breaks; every tie-break is kept.
#pragma once
#include "common.cuh"

namespace fseb {

__device__ __forceinline__ unsigned hibit0(unsigned v) { return v ? hibit(v) : 0u; }

__device__ inline unsigned d_min_tablelog(u64 srcSize, unsigned msv)
{
    unsigned const bySrc = hibit0((u32)srcSize) + 1, bySym = hibit0(msv) + 2;
    return bySrc < bySym ? bySrc : bySym;
}

__device__ inline unsigned d_optimal_tablelog(unsigned maxTableLog, u64 srcSize, unsigned msv, unsigned minus)
{
    unsigned const bySrc = hibit0((u32)(srcSize - 1)) - minus;
    unsigned const floorBits = d_min_tablelog(srcSize, msv);
    unsigned tl = maxTableLog ? maxTableLog : FSE_DEF_TLOG;
    if (bySrc < tl) tl = bySrc;
    if (floorBits > tl) tl = floorBits;
    if (tl < FSE_MIN_TLOG) tl = FSE_MIN_TLOG;
    if (tl > FSE_MAX_TLOG) tl = FSE_MAX_TLOG;
    return tl;
}

// ---- normalisation -------------------------------------------------------------------------
__device__ inline u64 d_normalize_fallback(short* norm, unsigned tl, const unsigned* count, u64 total, unsigned msv)
{
    const short PENDING = -2;
    u32 const lowThr = (u32)(total >> tl);
    u32 lowOne = (u32)((total * 3) >> (tl + 1));
    u32 given = 0, left;
    for (u32 s = 0; s <= msv; s++) {
        u32 const c = count[s];
        if (c == 0) { norm[s] = 0; continue; }
        if (c <= lowThr) { norm[s] = -1; given++; total -= c; continue; }
        if (c <= lowOne) { norm[s] = 1; given++; total -= c; continue; }
        norm[s] = PENDING;
    }
    left = (1u << tl) - given;
    if (left == 0) return 0;
    if ((total / left) > lowOne) {
        lowOne = (u32)((total * 3) / (left * 2));
        for (u32 s = 0; s <= msv; s++)
            if (norm[s] == PENDING && count[s] <= lowOne) { norm[s] = 1; given++; total -= count[s]; }
        left = (1u << tl) - given;
    }
    if (given == msv + 1) {
        u32 argmax = 0, vmax = 0;
        for (u32 s = 0; s <= msv; s++) if (count[s] > vmax) { vmax = count[s]; argmax = s; }
        norm[argmax] = (short)(norm[argmax] + (short)left);
        return 0;
    }
    if (total == 0) {
        for (u32 s = 0; left > 0; s = (s + 1) % (msv + 1))
            if (norm[s] > 0) { left--; norm[s]++; }
        return 0;
    }
    {   u64 const vlog = 62 - tl;
        u64 const mid = (1ULL << (vlog - 1)) - 1;
        u64 const rstep = (((1ULL << vlog) * left) + mid) / total;
        u64 run = mid;
        for (u32 s = 0; s <= msv; s++) {
            if (norm[s] != PENDING) continue;
            u64 const end = run + (u64)count[s] * rstep;
            u32 const w = (u32)(end >> vlog) - (u32)(run >> vlog);
            if (w < 1) return err(E_GENERIC);
            norm[s] = (short)w; run = end;
        }
    }
    return 0;
}

// returns tableLog, 0 (single symbol) or an error code
__device__ inline u64 d_normalize(short* norm, unsigned tl, const unsigned* count, u64 total, unsigned msv)
{
    const u32 roundUp[8] = { 0, 473195, 504333, 520860, 550000, 700000, 750000, 830000 };
    if (tl == 0) tl = FSE_DEF_TLOG;
    if (tl < FSE_MIN_TLOG) return err(E_GENERIC);
    if (tl > FSE_MAX_TLOG) return err(E_TLOG_TOO_LARGE);
    if (tl < d_min_tablelog(total, msv)) return err(E_GENERIC);
    u64 const scale = 62 - tl;
    u64 const step = (1ULL << 62) / total;
    u64 const vstep = 1ULL << (scale - 20);
    u32 const lowThr = (u32)(total >> tl);
    int toGive = 1 << tl;
    unsigned argmax = 0;
    short pmax = 0;
    for (u32 s = 0; s <= msv; s++) {
        u32 const c = count[s];
        if (c == total) return 0;
        if (c == 0) { norm[s] = 0; continue; }
        if (c <= lowThr) { norm[s] = -1; toGive--; continue; }
        u64 const scaled = (u64)c * step;
        short p = (short)(scaled >> scale);
        if (p < 8) p = (short)(p + ((scaled - ((u64)p << scale)) > vstep * roundUp[p]));
        if (p > pmax) { pmax = p; argmax = s; }
        norm[s] = p; toGive -= p;
    }
    if (-toGive >= (norm[argmax] >> 1)) {
        u64 const e = d_normalize_fallback(norm, tl, count, total, msv);
        if (is_err(e)) return e;
    } else norm[argmax] = (short)(norm[argmax] + (short)toGive);
    return tl;
}

// ---- NCount header -------------------------------------------------------------------------
__device__ inline u64 d_ncount_bound(unsigned msv, unsigned tl) { return msv ? (((u64)(msv + 1) * tl) >> 3) + 3 : 512; }

__device__ inline u64 d_write_ncount(u8* out, u64 cap, const short* norm, unsigned msv, unsigned tl)
{
    u64 o = 0;
    bool const guarded = cap < d_ncount_bound(msv, tl);
    unsigned const alphabet = msv + 1;
    if (tl > FSE_MAX_TLOG) return err(E_TLOG_TOO_LARGE);
    if (tl < FSE_MIN_TLOG) return err(E_GENERIC);
    u32 acc = tl - FSE_MIN_TLOG; int held = 4;
    int remaining = (1 << tl) + 1, threshold = 1 << tl, width = (int)tl + 1;
    unsigned sym = 0; bool afterZero = false;
#define FSEB_SPILL16() do { if (guarded && o + 2 > cap) return err(E_DST_TOO_SMALL); \
                            out[o] = (u8)acc; out[o + 1] = (u8)(acc >> 8); o += 2; acc >>= 16; } while (0)
    while (sym < alphabet && remaining > 1) {
        if (afterZero) {
            unsigned from = sym;
            while (sym < alphabet && !norm[sym]) sym++;
            if (sym == alphabet) break;
            while (sym >= from + 24) { from += 24; acc += 0xFFFFu << held; FSEB_SPILL16(); }
            while (sym >= from + 3) { from += 3; acc += 3u << held; held += 2; }
            acc += (sym - from) << held; held += 2;
            if (held > 16) { FSEB_SPILL16(); held -= 16; }
        }
        int v = norm[sym++];
        int const cut = (2 * threshold - 1) - remaining;
        remaining -= v < 0 ? -v : v;
        v++;
        if (v >= threshold) v += cut;
        acc += (u32)v << held;
        held += width; held -= (v < cut);
        afterZero = (v == 1);
        if (remaining < 1) return err(E_GENERIC);
        while (remaining < threshold) { width--; threshold >>= 1; }
        if (held > 16) { FSEB_SPILL16(); held -= 16; }
    }
    if (remaining != 1) return err(E_GENERIC);
    if (guarded && o + 2 > cap) return err(E_DST_TOO_SMALL);
    out[o] = (u8)acc; out[o + 1] = (u8)(acc >> 8);
    o += (u64)((held + 7) / 8);
#undef FSEB_SPILL16
    return o;
}

// Reads at most `hbSize` bytes of `in`.  *msvPtr is in/out (declared maximum -> last symbol present).
__device__ inline u64 d_read_ncount(short* norm, unsigned* msvPtr, unsigned* tlPtr, const u8* in, u64 hbSize)
{
    u8 pad[4] = { 0, 0, 0, 0 };
    bool const padded = hbSize < 4;
    u64 const trueSize = hbSize;
    if (padded) { for (u64 i = 0; i < hbSize; i++) pad[i] = in[i]; in = pad; hbSize = 4; }
    long long const end = (long long)hbSize;
    long long ip = 0;
    for (unsigned s = 0; s <= *msvPtr; s++) norm[s] = 0;
    u32 bits = rd32(in);
    int width = (int)(bits & 0xF) + (int)FSE_MIN_TLOG;
    if (width > (int)FSE_ABS_TLOG) return err(E_TLOG_TOO_LARGE);
    bits >>= 4; int held = 4;
    *tlPtr = (unsigned)width;
    int remaining = (1 << width) + 1, threshold = 1 << width; width++;
    unsigned sym = 0; bool afterZero = false;
    while ((remaining > 1) & (sym <= *msvPtr)) {
        if (afterZero) {
            unsigned upto = sym;
            while ((bits & 0xFFFF) == 0xFFFF) {
                upto += 24;
                if (ip < end - 5) { ip += 2; bits = rd32(in + ip) >> held; }
                else { bits >>= 16; held += 16; }
            }
            while ((bits & 3) == 3) { upto += 3; bits >>= 2; held += 2; }
            upto += bits & 3; held += 2;
            if (upto > *msvPtr) return err(E_MSV_TOO_SMALL);
            while (sym < upto) norm[sym++] = 0;
            if ((ip <= end - 7) || (ip + (held >> 3) <= end - 4)) { ip += held >> 3; held &= 7; bits = rd32(in + ip) >> held; }
            else bits >>= 2;
        }
        int const cut = (2 * threshold - 1) - remaining;
        int v;
        if ((bits & (u32)(threshold - 1)) < (u32)cut) { v = (int)(bits & (u32)(threshold - 1)); held += width - 1; }
        else { v = (int)(bits & (u32)(2 * threshold - 1)); if (v >= threshold) v -= cut; held += width; }
        v--;
        remaining -= v < 0 ? -v : v;
        norm[sym++] = (short)v;
        afterZero = !v;
        while (remaining < threshold) { width--; threshold >>= 1; }
        if ((ip <= end - 7) || (ip + (held >> 3) <= end - 4)) { ip += held >> 3; held &= 7; }
        else { held -= (int)(8 * (end - 4 - ip)); ip = end - 4; }
        bits = rd32(in + ip) >> (held & 31);
    }
    if (remaining != 1) return err(E_CORRUPT);
    if (held > 32) return err(E_CORRUPT);
    *msvPtr = sym - 1;
    ip += (held + 7) >> 3;
    if (padded && (u64)ip > trueSize) return err(E_CORRUPT);
    return (u64)ip;
}

// ---- symbol spreading (lib/fse_compress.c:96-122 == lib/fse_decompress.c:91-114) ------------
// One lane walks the table.  cellSym: u16[1<<tl].  Returns true iff the walk closes on cell 0.
__device__ inline bool d_spread_serial(u16* cellSym, const short* norm, unsigned msv, unsigned tl)
{
    u32 const size = 1u << tl, mask = size - 1;
    u32 const stride = (size >> 1) + (size >> 3) + 3;
    u32 high = size - 1, pos = 0;
    for (u32 s = 0; s <= msv; s++) if (norm[s] == -1) cellSym[high--] = (u16)s;
    for (u32 s = 0; s <= msv; s++) {
        int const n = norm[s];
        for (int k = 0; k < n; k++) {
            cellSym[pos] = (u16)s;
            do pos = (pos + stride) & mask; while (pos > high);
        }
    }
    return pos == 0;
}

// CTable image builder, one lane.  `start` scratch: u32[msv+2]; cellSym: u16[1<<tl].
// ct layout: u16 tableLog, u16 maxSV, u16 nextState[size], then {i32 deltaFindState,u32 deltaNbBits}[msv+1].
__device__ inline void d_build_ctable_serial(u32* ct, const short* norm, unsigned msv, unsigned tl, u16* cellSym, u32* start)
{
    u32 const size = 1u << tl;
    u16* const hdr = (u16*)ct;
    u16* const next = hdr + 2;
    u32* const tt = ct + 1 + (tl ? (size >> 1) : 1);
    hdr[0] = (u16)tl; hdr[1] = (u16)msv;
    start[0] = 0;
    for (u32 s = 1; s <= msv + 1; s++) start[s] = start[s - 1] + (u32)(norm[s - 1] == -1 ? 1 : norm[s - 1]);
    start[msv + 1] = size + 1;
    d_spread_serial(cellSym, norm, msv, tl);
    for (u32 u = 0; u < size; u++) next[start[cellSym[u]]++] = (u16)(size + u);
    u32 total = 0;
    for (u32 s = 0; s <= msv; s++) {
        int const n = norm[s];
        if (n == 0) { tt[2 * s + 1] = ((tl + 1) << 16) - size; tt[2 * s] = 0; continue; }   // (reference leaves deltaFindState unset)
        if (n == -1 || n == 1) { tt[2 * s + 1] = (tl << 16) - size; tt[2 * s] = total - 1; total++; }
        else {
            u32 const maxOut = tl - hibit((u32)n - 1);
            tt[2 * s + 1] = (maxOut << 16) - ((u32)n << maxOut);
            tt[2 * s] = total - (u32)n; total += (u32)n;
        }
    }
}

// DTable image builder, one lane.  wide=false: {u16 newState,u8 symbol,u8 nbBits}; wide=true (U16):
// {u16 newState, nbBits:4, symbol:12}.  nextOf scratch: u16[msv+1]; cellSym: u16[1<<tl].
template <bool WIDE>
__device__ inline u64 d_build_dtable_serial(u32* dt, const short* norm, unsigned msv, unsigned tl,
                                            unsigned msvLimit, unsigned tlLimit, u16* cellSym, u16* nextOf)
{
    u32 const size = 1u << tl;
    unsigned fast = 1;
    if (msv > msvLimit) return err(E_MSV_TOO_LARGE);
    if (tl > tlLimit) return err(E_TLOG_TOO_LARGE);
    for (u32 s = 0; s <= msv; s++) {
        if (norm[s] == -1) nextOf[s] = 1;
        else { if (norm[s] >= (short)(1 << (tl - 1))) fast = 0; nextOf[s] = (u16)norm[s]; }
    }
    dt[0] = tl | (fast << 16);
    if (!d_spread_serial(cellSym, norm, msv, tl)) return err(E_GENERIC);
    for (u32 u = 0; u < size; u++) {
        u32 const sym = cellSym[u];
        u32 const x = nextOf[sym]++;
        u32 const nb = tl - hibit(x);
        u32 const ns = ((x << nb) - size) & 0xFFFF;
        dt[1 + u] = WIDE ? (ns | (nb << 16) | (sym << 20)) : (ns | (sym << 16) | (nb << 24));
    }
    return 0;
}

}  // namespace fseb
