// huf_dev.cuh -- device-side Huff0 header parsing shared by the decode kernels.
//
//   HUF_readStats   lib/entropy_common.c:154-215  (raw-nibble or FSE-compressed weights)
// One lane parses one header; all arithmetic and every rejection rule is kept, so a malformed
// header yields the same error code as the CPU library.
#pragma once
#include "common.cuh"
#include "fse_dev.cuh"
#include "bitsrc_dev.cuh"

namespace fseb {

// HUF_selectDecoder (lib/huf_decompress.c:1001-1051): 1 = the reference decodes this block with its double-symbol decoder.
__device__ inline u32 d_huf_select_decoder(u64 dstSize, u64 cSrcSize)
{
    const u16 cost[16][4] = {
        {0, 0, 1, 1}, {0, 0, 1, 1}, {38, 130, 1313, 74}, {448, 128, 1353, 74}, {556, 128, 1353, 74},
        {714, 128, 1418, 74}, {883, 128, 1437, 74}, {897, 128, 1515, 75}, {926, 128, 1613, 75},
        {947, 128, 1729, 77}, {1107, 128, 2083, 81}, {1177, 128, 2379, 87}, {1242, 128, 2415, 93},
        {1349, 128, 2644, 106}, {1455, 128, 2422, 124}, {722, 128, 1891, 145} };
    u32 const q = (cSrcSize >= dstSize) ? 15u : (u32)(cSrcSize * 16 / dstSize);
    u32 const d256 = (u32)(dstSize >> 8);
    u32 const t0 = cost[q][0] + cost[q][1] * d256;
    u32 t1 = cost[q][2] + cost[q][3] * d256;
    t1 += t1 >> 3;
    return t1 < t0;
}

// The FSE-compressed form of the weights (entropy_common.c:178-183, FSE_decompress with tableLog <= 6): in[1 .. iSize] -> weights,
// returns their number or an error.
__device__ inline u64 d_huf_read_weights_fse(u8* weights, u64 hwSize, const u8* in, u64 iSize)
{
    short norm[FSE_MAX_SV + 1];
    u32 dt[1 + 64];
    u16 cellSym[64];
    u16 nextOf[FSE_MAX_SV + 1];
    unsigned tl = 0, msv = FSE_MAX_SV;
    u64 const h = d_read_ncount(norm, &msv, &tl, in + 1, iSize);
    if (is_err(h)) return h;
    if (tl > 6) return err(E_TLOG_TOO_LARGE);
    u64 const r = d_build_dtable_serial<false>(dt, norm, msv, tl, FSE_MAX_SV, FSE_MAX_TLOG, cellSym, nextOf);
    if (is_err(r)) return r;
    return d_fse_decode_serial(weights, hwSize - 1, in + 1 + h, iSize - h, dt);
}

// weights: u8[hwSize] (hwSize = 256), rankStats: u32[13].  Returns header bytes consumed or an error.
__device__ inline u64 d_huf_read_stats(u8* weights, u64 hwSize, u32* rankStats, u32* nbSymPtr, u32* tlPtr,
                                       const u8* in, u64 srcSize)
{
    if (!srcSize) return err(E_SRC_WRONG);
    u64 iSize = in[0], oSize;
    if (iSize >= 128) {                                   // raw 4-bit weights (entropy_common.c:167-177)
        oSize = iSize - 127;
        iSize = (oSize + 1) / 2;
        if (iSize + 1 > srcSize) return err(E_SRC_WRONG);
        if (oSize >= hwSize) return err(E_CORRUPT);
        for (u64 n = 0; n < oSize; n += 2) {
            u8 const v = in[1 + n / 2];
            weights[n] = v >> 4; weights[n + 1] = v & 15;
        }
    } else {                                              // FSE-compressed weights, tableLog <= 6 (:178-183)
        if (iSize + 1 > srcSize) return err(E_SRC_WRONG);
        oSize = d_huf_read_weights_fse(weights, hwSize, in, iSize);
        if (is_err(oSize)) return oSize;
    }
    for (unsigned i = 0; i <= HUF_MAX_TLOG; i++) rankStats[i] = 0;
    u32 total = 0;
    for (u64 n = 0; n < oSize; n++) {
        if (weights[n] >= HUF_MAX_TLOG) return err(E_CORRUPT);
        rankStats[weights[n]]++;
        total += (1u << weights[n]) >> 1;
    }
    if (total == 0) return err(E_CORRUPT);
    u32 const tl = hibit(total) + 1;
    if (tl > HUF_MAX_TLOG) return err(E_CORRUPT);
    *tlPtr = tl;
    u32 const rest = (1u << tl) - total;
    u32 const lastW = hibit(rest) + 1;
    if ((1u << hibit(rest)) != rest) return err(E_CORRUPT);
    weights[oSize] = (u8)lastW;
    rankStats[lastW]++;
    if ((rankStats[1] < 2) || (rankStats[1] & 1)) return err(E_CORRUPT);
    *nbSymPtr = (u32)(oSize + 1);
    return iSize + 1;
}

}  // namespace fseb
