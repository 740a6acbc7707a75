// eg_project_vjp.cuh -- the projection backward of one Gaussian (gsplat==1.0.0 fully_fused_projection bwd,
// SURVEY.md Appendix A.6, behind /root/reference/edgegaussians/models/edge_gs.py:250-268) preceded by the VJP
// of rendering.py's  opacities * compensations  and followed by ExpBackward / SigmoidBackward of
// edge_gs.py:253-254 when RAW.  Shared by eg_project_bwd (gsplat-shaped autograd path) and eg_splat_bwd
// (fused Gaussian-major backward).
#pragma once
#include "eg_common.cuh"

// g0 = (v_mean2d.x, v_mean2d.y, absgrad.x, absgrad.y), g1 = (v_conic.a, v_conic.b, v_conic.c, v_opacity_eff);
// r1 = (conic a, b, c, comp); (mx,my,mz), q4 (w,x,y,z), s[3], o are the parameters AS GIVEN (raw when RAW).
template <bool RAW>
__device__ __forceinline__ void eg_project_vjp(const eg_config &cfg, const EgCam &cam, const float4 r1,
                                               const float4 g0, const float4 g1, const float mx, const float my,
                                               const float mz, const float4 q4, float (&s)[3], float o,
                                               const float v_depth, float (&vm)[3], float (&vs)[3], float (&vq)[4],
                                               float &vo) {
    const float *R = cam.R;
    const float A = r1.x, B = r1.y, C = r1.z, comp = r1.w;
    if (RAW) o = 1.0f / (1.0f + expf(-o));
    float v_comp = 0.0f;
    if (cfg.antialiased) {
        vo = g1.w * comp;
        v_comp = g1.w * o;
    } else {
        vo = g1.w;
    }
    if (RAW) vo *= o * (1.0f - o);

    // v_Sigma2 = -Cn V Cn  (+ blur compensation term)
    const float vA = g1.x, vB = 0.5f * g1.y, vC = g1.z;
    const float cv00 = A * vA + B * vB, cv01 = A * vB + B * vC;
    const float cv10 = B * vA + C * vB, cv11 = B * vB + C * vC;
    float s00 = -(cv00 * A + cv01 * B), s01 = -(cv00 * B + cv01 * C);
    float s10 = -(cv10 * A + cv11 * B), s11 = -(cv10 * B + cv11 * C);
    if (cfg.antialiased) {
        const float detc = A * C - B * B;
        const float v_sqr = v_comp * 0.5f / (comp + 1e-6f);
        const float om = 1.0f - comp * comp;
        s00 += v_sqr * (om * A - cfg.eps2d * detc);
        s01 += v_sqr * (om * B);
        s10 += v_sqr * (om * B);
        s11 += v_sqr * (om * C - cfg.eps2d * detc);
    }

    // recompute the forward intermediates
    const float x = R[0] * mx + R[1] * my + R[2] * mz + cam.t[0];
    const float y = R[3] * mx + R[4] * my + R[5] * mz + cam.t[1];
    const float z = R[6] * mx + R[7] * my + R[8] * mz + cam.t[2];
    const float qn = sqrtf(q4.x * q4.x + q4.y * q4.y + q4.z * q4.z + q4.w * q4.w);
    const float iqn = 1.0f / qn;
    const float qw = q4.x * iqn, qx = q4.y * iqn, qy = q4.z * iqn, qz = q4.w * iqn;
    float Rq[3][3];
    Rq[0][0] = 1.f - 2.f * (qy * qy + qz * qz); Rq[0][1] = 2.f * (qx * qy - qw * qz); Rq[0][2] = 2.f * (qx * qz + qw * qy);
    Rq[1][0] = 2.f * (qx * qy + qw * qz); Rq[1][1] = 1.f - 2.f * (qx * qx + qz * qz); Rq[1][2] = 2.f * (qy * qz - qw * qx);
    Rq[2][0] = 2.f * (qx * qz - qw * qy); Rq[2][1] = 2.f * (qy * qz + qw * qx); Rq[2][2] = 1.f - 2.f * (qx * qx + qy * qy);
    if (RAW) { s[0] = expf(s[0]); s[1] = expf(s[1]); s[2] = expf(s[2]); }
    float M[3][3], S[3][3], Tm[3][3], Sc[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) M[i][j] = Rq[i][j] * s[j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) S[i][j] = M[i][0] * M[j][0] + M[i][1] * M[j][1] + M[i][2] * M[j][2];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Tm[i][j] = R[3 * i] * S[0][j] + R[3 * i + 1] * S[1][j] + R[3 * i + 2] * S[2][j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Sc[i][j] = Tm[i][0] * R[3 * j] + Tm[i][1] * R[3 * j + 1] + Tm[i][2] * R[3 * j + 2];

    const float fx = cam.fx, fy = cam.fy;
    const float lim_x = 1.3f * (0.5f * (float)cfg.width / fx), lim_y = 1.3f * (0.5f * (float)cfg.height / fy);
    const float rz = 1.0f / z, rz2 = rz * rz, rz3 = rz2 * rz;
    const float xr = x * rz, yr = y * rz;
    const float tx = z * fminf(lim_x, fmaxf(-lim_x, xr));
    const float ty = z * fminf(lim_y, fmaxf(-lim_y, yr));
    // J = [[j00, 0, j02], [0, j11, j12]]
    const float j00 = fx * rz, j11 = fy * rz, j02 = -fx * tx * rz2, j12 = -fy * ty * rz2;
    const float J[2][3] = {{j00, 0.f, j02}, {0.f, j11, j12}};
    const float V2[2][2] = {{s00, s01}, {s10, s11}};
    // v_Sigma_c = J^T v_S2 J
    float vSc[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            vSc[i][j] = J[0][i] * (V2[0][0] * J[0][j] + V2[0][1] * J[1][j]) + J[1][i] * (V2[1][0] * J[0][j] + V2[1][1] * J[1][j]);
    // v_J = v_S2 J Sc^T + v_S2^T J Sc
    float JSct[2][3], JSc[2][3];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            JSct[a][j] = J[a][0] * Sc[j][0] + J[a][1] * Sc[j][1] + J[a][2] * Sc[j][2];
            JSc[a][j] = J[a][0] * Sc[0][j] + J[a][1] * Sc[1][j] + J[a][2] * Sc[2][j];
        }
    float vJ[2][3];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            vJ[a][j] = V2[a][0] * JSct[0][j] + V2[a][1] * JSct[1][j] + V2[0][a] * JSc[0][j] + V2[1][a] * JSc[1][j];

    const float vmx = g0.x, vmy = g0.y;
    float vp0 = fx * rz * vmx, vp1 = fy * rz * vmy;
    float vp2 = -(fx * x * vmx + fy * y * vmy) * rz2;
    if (xr <= lim_x && xr >= -lim_x) vp0 += -fx * rz2 * vJ[0][2];
    else vp2 += -fx * rz3 * vJ[0][2] * tx;
    if (yr <= lim_y && yr >= -lim_y) vp1 += -fy * rz2 * vJ[1][2];
    else vp2 += -fy * rz3 * vJ[1][2] * ty;
    vp2 += -fx * rz2 * vJ[0][0] - fy * rz2 * vJ[1][1] + 2.0f * fx * tx * rz3 * vJ[0][2] + 2.0f * fy * ty * rz3 * vJ[1][2];
    vp2 += v_depth;
#pragma unroll
    for (int c = 0; c < 3; ++c) vm[c] = R[c] * vp0 + R[3 + c] * vp1 + R[6 + c] * vp2;

    // v_Sigma = R^T v_Sc R ; v_M = (v_S + v_S^T) M
    float RtV[3][3], vS[3][3], vM[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) RtV[i][j] = R[i] * vSc[0][j] + R[3 + i] * vSc[1][j] + R[6 + i] * vSc[2][j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) vS[i][j] = RtV[i][0] * R[j] + RtV[i][1] * R[3 + j] + RtV[i][2] * R[6 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            vM[i][j] = (vS[i][0] + vS[0][i]) * M[0][j] + (vS[i][1] + vS[1][i]) * M[1][j] + (vS[i][2] + vS[2][i]) * M[2][j];
    float vRq[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        vs[j] = Rq[0][j] * vM[0][j] + Rq[1][j] * vM[1][j] + Rq[2][j] * vM[2][j];
        if (RAW) vs[j] *= s[j];
#pragma unroll
        for (int i = 0; i < 3; ++i) vRq[i][j] = vM[i][j] * s[j];
    }
    float vqh[4];
    vqh[0] = 2.f * (qx * (vRq[2][1] - vRq[1][2]) + qy * (vRq[0][2] - vRq[2][0]) + qz * (vRq[1][0] - vRq[0][1]));
    vqh[1] = 2.f * (-2.f * qx * (vRq[1][1] + vRq[2][2]) + qy * (vRq[0][1] + vRq[1][0]) + qz * (vRq[0][2] + vRq[2][0]) + qw * (vRq[2][1] - vRq[1][2]));
    vqh[2] = 2.f * (qx * (vRq[0][1] + vRq[1][0]) - 2.f * qy * (vRq[0][0] + vRq[2][2]) + qz * (vRq[1][2] + vRq[2][1]) + qw * (vRq[0][2] - vRq[2][0]));
    vqh[3] = 2.f * (qx * (vRq[0][2] + vRq[2][0]) + qy * (vRq[1][2] + vRq[2][1]) - 2.f * qz * (vRq[0][0] + vRq[1][1]) + qw * (vRq[1][0] - vRq[0][1]));
    const float dotq = vqh[0] * qw + vqh[1] * qx + vqh[2] * qy + vqh[3] * qz;
    vq[0] = (vqh[0] - dotq * qw) * iqn;
    vq[1] = (vqh[1] - dotq * qx) * iqn;
    vq[2] = (vqh[2] - dotq * qy) * iqn;
    vq[3] = (vqh[3] - dotq * qz) * iqn;
}
