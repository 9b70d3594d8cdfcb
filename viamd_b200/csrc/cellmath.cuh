// cellmath.cuh — per-point arithmetic of the cell-list build shared by cells.cu (static index lists) and within.cu (per-frame lists).
#pragma once
#include "common.cuh"

namespace mdg {

// ---------------------------------------------------------------------------------------------------------------
// Point binning. vec4_linear_combine_3(r - origin, I) (core/md_vec_math.h:1323): ((I0*a.x) + (I1*a.y)) + (I2*a.z).
// ---------------------------------------------------------------------------------------------------------------
MDG_D void cart_to_fract(float s[3], const float r[3], const FrameGeom& g) {
    const float ax = __fsub_rn(r[0], g.origin[0]), ay = __fsub_rn(r[1], g.origin[1]), az = __fsub_rn(r[2], g.origin[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = __fmul_rn(g.I[0][k], ax);
        v = __fadd_rn(v, __fmul_rn(g.I[1][k], ay));
        v = __fadd_rn(v, __fmul_rn(g.I[2][k], az));
        s[k] = v;
    }
}

}  // namespace mdg
