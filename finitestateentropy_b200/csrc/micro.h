// micro.h -- op codes of the single-CTA table kernels (micro.cu) shared with the C-ABI layer (capi.cu).
#pragma once
#include <stdint.h>
namespace fseb {
struct MicroArgs { unsigned long long a[6]; };
enum {
    MOP_NORMALIZE = 1, MOP_WRITE_NCOUNT, MOP_READ_NCOUNT, MOP_BUILD_CTABLE, MOP_BUILD_DTABLE,
    MOP_HUF_BUILD_CTABLE, MOP_HUF_WRITE_CTABLE, MOP_HUF_READ_STATS, MOP_HUF_READ_DTABLE_X1,
    MOP_FSE_ENCODE_CT, MOP_FSE_DECODE_DT, MOP_HUF_ENCODE4X_CT, MOP_HUF_DECODE4X1_DT,
    MOP_HUF_ENCODE1X_CT, MOP_HUF_DECODE1X1_DT, MOP_HUF_READ_DTABLE_X2, MOP_HUF_DECODE4X2_DT, MOP_HUF_DECODE1X2_DT
};
}
