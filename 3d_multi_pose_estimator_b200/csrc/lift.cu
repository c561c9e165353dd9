// Stage 3: per-person MLP-input encoder with pairwise-DLT hint, and the triangulation baseline.
//
// Reference behaviour restated (paths relative to the reference root):
//   utils/pose_estimator_dataset_from_json.py:63-101   get_3D_from_triangulation (joint id > 0 only,
//        every camera pair, cv2.undistortPoints + cv2.triangulatePoints, plain mean over pairs)
//   utils/pose_estimator_dataset_from_json.py:237-289  PoseEstimatorDataset dict branch (14 numbers per
//        joint per camera; triangulation slots written into EVERY used-camera block)
//   utils/pose_estimator_utils.py:52-75                triangulate(): pairwise DLT, upper median on one
//        coordinate, keep pairs within 0.05 m, mean
//   OpenCV: undistortPoints = 5 fixed-point iterations of the inverse Brown model in fp64;
//           triangulatePoints = right singular vector of the smallest singular value of the 4x4 DLT matrix.
//
// One warp per person, one lane per joint. Everything geometric stays in fp64 registers: the 4x4
// null-vector problem is solved by a register-resident one-sided (Hestenes) Jacobi SVD, which works on
// A directly (no A^T A, so the condition number is not squared - the narrow-baseline ARP stereo pair
// needs that).
#include "common.cuh"

namespace b200pose {

constexpr int kMaxCams = B200POSE_MAX_CAMERAS;
constexpr int kJ = B200POSE_N_JOINTS;

struct LiftTables {
    int n_cameras, v_pe;
    float W, Hh;
    const int* pe_slot;
    const float* t_cam2root32;
    const double* k64;
    const double* dist64;
    const double* p64;
};

// cv2.undistortPoints(pt, K, dist): normalised coordinates after 5 iterations
__device__ __forceinline__ void undistort(double u, double v, const double* __restrict__ k4, const double* __restrict__ d5,
                                          double& xo, double& yo)
{
    const double fx = k4[0], fy = k4[1], cx = k4[2], cy = k4[3];
    const double k1 = d5[0], k2 = d5[1], p1 = d5[2], p2 = d5[3], k3 = d5[4];
    double x = (u - cx) * (1.0 / fx);
    double y = (v - cy) * (1.0 / fy);
    const double x0 = x, y0 = y;
#pragma unroll 1
    for (int it = 0; it < 5; ++it) {
        const double r2 = x * x + y * y;
        const double icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2);
        if (icdist < 0) { x = x0; y = y0; break; }
        const double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
        const double dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
        x = (x0 - dx) * icdist;
        y = (y0 - dy) * icdist;
    }
    xo = x; yo = y;
}

// cv2.triangulatePoints for one point and two views, dehomogenised. P1,P2: 3x4 row-major fp64.
__device__ void triangulate_pair(const double* __restrict__ P1, const double* __restrict__ P2,
                                 double x1, double y1, double x2, double y2, double (&X)[3])
{
    double A[4][4], V[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        A[0][k] = x1 * P1[8 + k] - P1[k];
        A[1][k] = y1 * P1[8 + k] - P1[4 + k];
        A[2][k] = x2 * P2[8 + k] - P2[k];
        A[3][k] = y2 * P2[8 + k] - P2[4 + k];
#pragma unroll
        for (int i = 0; i < 4; ++i) V[i][k] = (i == k) ? 1.0 : 0.0;
    }
    // one-sided Jacobi: orthogonalise the columns of A, accumulate the rotations in V
#pragma unroll 1
    for (int sweep = 0; sweep < 16; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                double al = 0, be = 0, ga = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) { al += A[i][p] * A[i][p]; be += A[i][q] * A[i][q]; ga += A[i][p] * A[i][q]; }
                if (fabs(ga) > 1e-16 * sqrt(al * be) && ga != 0.0) {
                    rotated = true;
                    const double zeta = (be - al) / (2.0 * ga);
                    const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double ap = A[i][p], aq = A[i][q];
                        A[i][p] = c * ap - s * aq; A[i][q] = s * ap + c * aq;
                        const double vp = V[i][p], vq = V[i][q];
                        V[i][p] = c * vp - s * vq; V[i][q] = s * vp + c * vq;
                    }
                }
            }
        }
        if (!rotated) break;
    }
    int best = 0; double bn = 1e300;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double n = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) n += A[i][q] * A[i][q];
        if (n < bn) { bn = n; best = q; }
    }
    double v0 = 0, v1 = 0, v2 = 0, v3 = 1;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (q == best) { v0 = V[0][q]; v1 = V[1][q]; v2 = V[2][q]; v3 = V[3][q]; }
    X[0] = v0 / v3; X[1] = v1 / v3; X[2] = v2 / v3;
}

// ---- MLP-input encoder -------------------------------------------------------------------------
__global__ void __launch_bounds__(128) encode_persons_kernel(
    int n_persons, const int* __restrict__ person_sk, const double* __restrict__ sk_xy, const float* __restrict__ sk_vp,
    const uint32_t* __restrict__ sk_mask, LiftTables t,
    float* __restrict__ x_f32, int ld_f32, __nv_bfloat16* __restrict__ x_hi, __nv_bfloat16* __restrict__ x_lo, int ld_planes,
    uint8_t* __restrict__ valid)
{
    const int person = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (person >= n_persons) return;
    const int C = t.n_cameras;
    // cameras of this person in used_cameras order (dataset.py raw_input is built in that order,
    // test/metrics_from_model.py:248-252)
    int cam_of_slot[kMaxCams];
    for (int s = 0; s < t.v_pe; ++s) cam_of_slot[s] = -1;
    for (int c = 0; c < C; ++c) { const int s = t.pe_slot[c]; if (s >= 0) cam_of_slot[s] = c; }

    const int j = lane;
    float abs_sum = 0.f;
    double ux[kMaxCams], uy[kMaxCams];
    uint32_t have = 0;                                    // slots where joint j is present
    const int row_len = kJ * 14 * t.v_pe;
    if (j < kJ) {
        for (int s = 0; s < t.v_pe; ++s) {
            const int c = cam_of_slot[s];
            const int sk = (c >= 0) ? person_sk[(size_t)person * C + c] : -1;
            float o[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) o[i] = 0.f;
            if (sk >= 0 && ((sk_mask[sk] >> j) & 1u)) {
                const double x = sk_xy[(size_t)sk * 36 + 2 * j], y = sk_xy[(size_t)sk * 36 + 2 * j + 1];
                double nx, ny;
                undistort(x, y, t.k64 + 4 * c, t.dist64 + 5 * c, nx, ny);
                ux[s] = nx; uy[s] = ny;
                have |= 1u << s;
                const double w2 = (double)t.W / 2.0, h2 = (double)t.Hh / 2.0;
                o[0] = sk_vp[(size_t)sk * 36 + 2 * j];
                o[1] = __double2float_rn((x - w2) / w2);
                o[2] = __double2float_rn((y - h2) / h2);
                o[3] = sk_vp[(size_t)sk * 36 + 2 * j + 1];
                const float* T = t.t_cam2root32 + 16 * c;
                const float fx_ = __double2float_rn(nx), fy_ = __double2float_rn(ny);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    o[4 + i] = __fdiv_rn(T[4 * i + 3], 10.0f);
                    float acc = __fmul_rn(T[4 * i], fx_);
                    acc = __fmaf_rn(T[4 * i + 1], fy_, acc);
                    acc = __fmaf_rn(T[4 * i + 2], 1.0f, acc);
                    acc = __fmaf_rn(T[4 * i + 3], 0.0f, acc);
                    o[7 + i] = __fdiv_rn(acc, 10.0f);
                }
            }
            const size_t base = (size_t)s * (kJ * 14) + (size_t)j * 14;
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                abs_sum += fabsf(o[i]);
                if (x_f32) x_f32[(size_t)person * ld_f32 + base + i] = o[i];
                if (x_hi) {
                    __nv_bfloat16 h, l;
                    split_bf16(o[i], h, l);
                    x_hi[(size_t)person * ld_planes + base + i] = h;
                    x_lo[(size_t)person * ld_planes + base + i] = l;
                }
            }
        }
        // pairwise triangulation hint (joint id > 0: dataset.py:75)
        float tri[4] = {0.f, 0.f, 0.f, 0.f};
        if (j > 0 && __popc(have) >= 2) {
            double acc[3] = {0, 0, 0};
            int n = 0;
            for (int s1 = 0; s1 < t.v_pe; ++s1) {
                if (!((have >> s1) & 1u)) continue;
                for (int s2 = s1 + 1; s2 < t.v_pe; ++s2) {
                    if (!((have >> s2) & 1u)) continue;
                    double X[3];
                    triangulate_pair(t.p64 + 12 * cam_of_slot[s1], t.p64 + 12 * cam_of_slot[s2], ux[s1], uy[s1], ux[s2], uy[s2], X);
                    acc[0] += X[0]; acc[1] += X[1]; acc[2] += X[2];
                    ++n;
                }
            }
            tri[0] = 1.0f;
#pragma unroll
            for (int i = 0; i < 3; ++i) tri[1 + i] = __double2float_rn((acc[i] / (double)n) / 10.0);
        }
        for (int s = 0; s < t.v_pe; ++s) {                 // every used-camera block (dataset.py:280-285)
            const size_t base = (size_t)s * (kJ * 14) + (size_t)j * 14 + 10;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                abs_sum += fabsf(tri[i]);
                if (x_f32) x_f32[(size_t)person * ld_f32 + base + i] = tri[i];
                if (x_hi) {
                    __nv_bfloat16 h, l;
                    split_bf16(tri[i], h, l);
                    x_hi[(size_t)person * ld_planes + base + i] = h;
                    x_lo[(size_t)person * ld_planes + base + i] = l;
                }
            }
        }
    }
    // zero the K padding of the planes
    if (x_hi) {
        for (int c = row_len + lane; c < ld_planes; c += 32) {
            x_hi[(size_t)person * ld_planes + c] = __float2bfloat16_rn(0.f);
            x_lo[(size_t)person * ld_planes + c] = __float2bfloat16_rn(0.f);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) abs_sum += __shfl_xor_sync(0xffffffffu, abs_sum, o);
    if (valid && lane == 0) valid[person] = abs_sum > 1.0f ? 1 : 0;          // dataset.py:287
}

// ---- triangulation baseline ---------------------------------------------------------------------
__global__ void __launch_bounds__(128) triangulate_kernel(
    int n_persons, const int* __restrict__ person_sk, const double* __restrict__ sk_xy, const uint32_t* __restrict__ sk_mask,
    LiftTables t, int median_axis, double* __restrict__ xyz, uint8_t* __restrict__ mask)
{
    const int person = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int j = threadIdx.x & 31;
    if (person >= n_persons || j >= kJ) return;
    const int C = t.n_cameras;
    double ux[kMaxCams], uy[kMaxCams];
    uint32_t have = 0;
    for (int c = 0; c < C; ++c) {                           // camera index order (metrics_from_triangulation.py:239)
        const int sk = person_sk[(size_t)person * C + c];
        if (sk >= 0 && ((sk_mask[sk] >> j) & 1u)) {
            undistort(sk_xy[(size_t)sk * 36 + 2 * j], sk_xy[(size_t)sk * 36 + 2 * j + 1], t.k64 + 4 * c, t.dist64 + 5 * c, ux[c], uy[c]);
            have |= 1u << c;
        }
    }
    double* out = xyz + ((size_t)person * kJ + j) * 3;
    const int ncam = __popc(have);
    if (ncam < 2) {
        out[0] = out[1] = out[2] = 0.0;
        mask[(size_t)person * kJ + j] = 0;
        return;
    }
    // pass 1: the coordinate used by the median filter, for every pair (<= 496 pairs; kept implicit)
    // The pair results are recomputed in pass 2 instead of being stored: registers over local memory.
    const int npairs = ncam * (ncam - 1) / 2;
    const int target = npairs / 2;                         // index of the upper median in sorted order
    // selection by counting: median = value v with exactly `target` values smaller (ties by pair order)
    double med = 0.0;
    {
        // gather the median-axis values in local memory (npairs <= 496)
        double vals[64];
        const bool small = npairs <= 64;
        int n = 0;
        for (int c1 = 0; c1 < C; ++c1) {
            if (!((have >> c1) & 1u)) continue;
            for (int c2 = c1 + 1; c2 < C; ++c2) {
                if (!((have >> c2) & 1u)) continue;
                double X[3];
                triangulate_pair(t.p64 + 12 * c1, t.p64 + 12 * c2, ux[c1], uy[c1], ux[c2], uy[c2], X);
                const double d = median_axis == 0 ? X[0] : (median_axis == 1 ? X[1] : X[2]);
                if (small) vals[n] = d;
                ++n;
            }
        }
        if (small) {
            for (int i = 0; i < npairs; ++i) {
                int less = 0, eq_before = 0;
                for (int k = 0; k < npairs; ++k) {
                    less += vals[k] < vals[i];
                    eq_before += (vals[k] == vals[i]) && (k < i);
                }
                if (less + eq_before == target) med = vals[i];
            }
        } else {
            // more than 64 pairs (> 11 cameras seeing the joint): selection by repeated recomputation
            double lo = -1e300;
            int taken = 0;
            while (true) {
                double best = 1e300; int cnt = 0;
                for (int c1 = 0; c1 < C; ++c1) {
                    if (!((have >> c1) & 1u)) continue;
                    for (int c2 = c1 + 1; c2 < C; ++c2) {
                        if (!((have >> c2) & 1u)) continue;
                        double X[3];
                        triangulate_pair(t.p64 + 12 * c1, t.p64 + 12 * c2, ux[c1], uy[c1], ux[c2], uy[c2], X);
                        const double d = median_axis == 0 ? X[0] : (median_axis == 1 ? X[1] : X[2]);
                        if (d > lo) { if (d < best) { best = d; cnt = 1; } else if (d == best) ++cnt; }
                    }
                }
                if (taken + cnt > target) { med = best; break; }
                taken += cnt; lo = best;
            }
        }
    }
    double acc[3] = {0, 0, 0};
    int kept = 0;
    for (int c1 = 0; c1 < C; ++c1) {
        if (!((have >> c1) & 1u)) continue;
        for (int c2 = c1 + 1; c2 < C; ++c2) {
            if (!((have >> c2) & 1u)) continue;
            double X[3];
            triangulate_pair(t.p64 + 12 * c1, t.p64 + 12 * c2, ux[c1], uy[c1], ux[c2], uy[c2], X);
            const double d = median_axis == 0 ? X[0] : (median_axis == 1 ? X[1] : X[2]);
            if (fabs(d - med) < 0.05) { acc[0] += X[0]; acc[1] += X[1]; acc[2] += X[2]; ++kept; }
        }
    }
    out[0] = acc[0] / kept; out[1] = acc[1] / kept; out[2] = acc[2] / kept;
    mask[(size_t)person * kJ + j] = 1;
}

}  // namespace b200pose

using namespace b200pose;

static LiftTables make_tables(const b200pose_cameras* cams) {
    LiftTables t;
    t.n_cameras = cams->n_cameras; t.v_pe = cams->v_pe; t.W = cams->image_width; t.Hh = cams->image_height;
    t.pe_slot = cams->pe_slot; t.t_cam2root32 = cams->t_cam2root32; t.k64 = cams->k64; t.dist64 = cams->dist64; t.p64 = cams->p64;
    return t;
}

extern "C" __attribute__((visibility("default"))) int b200pose_encode_persons(int32_t n_persons, const int32_t* person_sk, const double* sk_xy, const float* sk_vp,
                                       const uint32_t* sk_mask, const b200pose_cameras* cams,
                                       float* x_f32, int32_t ld_f32, uint16_t* x_hi, uint16_t* x_lo, int32_t ld_planes,
                                       uint8_t* valid, void* stream)
{
    B2_CHECK_ARG(person_sk && sk_xy && sk_vp && sk_mask && cams, "encode_persons: null input");
    B2_CHECK_ARG(cams->n_cameras <= B200POSE_MAX_CAMERAS, "encode_persons: too many cameras");
    B2_CHECK_ARG((x_hi == nullptr) == (x_lo == nullptr), "encode_persons: planes go together");
    const int row_len = B200POSE_N_JOINTS * 14 * cams->v_pe;
    if (x_f32) B2_CHECK_ARG(ld_f32 >= row_len, "encode_persons: ld_f32 too small");
    if (x_hi) B2_CHECK_ARG(ld_planes % 64 == 0 && ld_planes >= row_len, "encode_persons: bad ld_planes");
    if (n_persons == 0) return B200POSE_OK;
    encode_persons_kernel<<<ceil_div(n_persons, 4), 128, 0, (cudaStream_t)stream>>>(
        n_persons, person_sk, sk_xy, sk_vp, sk_mask, make_tables(cams), x_f32, ld_f32,
        reinterpret_cast<__nv_bfloat16*>(x_hi), reinterpret_cast<__nv_bfloat16*>(x_lo), ld_planes, valid);
    B2_CHECK_LAUNCH();
    return B200POSE_OK;
}

extern "C" __attribute__((visibility("default"))) int b200pose_triangulate(int32_t n_persons, const int32_t* person_sk, const double* sk_xy,
                                    const uint32_t* sk_mask, const b200pose_cameras* cams, int32_t median_axis,
                                    double* xyz, uint8_t* mask, void* stream)
{
    B2_CHECK_ARG(person_sk && sk_xy && sk_mask && cams && xyz && mask, "triangulate: null pointer");
    B2_CHECK_ARG(median_axis >= 0 && median_axis < 3, "triangulate: median_axis must be 0..2");
    B2_CHECK_ARG(cams->n_cameras <= B200POSE_MAX_CAMERAS, "triangulate: too many cameras");
    if (n_persons == 0) return B200POSE_OK;
    triangulate_kernel<<<ceil_div(n_persons, 4), 128, 0, (cudaStream_t)stream>>>(
        n_persons, person_sk, sk_xy, sk_mask, make_tables(cams), median_axis, xyz, mask);
    B2_CHECK_LAUNCH();
    return B200POSE_OK;
}
