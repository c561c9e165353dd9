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
// One CTA per person; threads take (camera, joint) undistortions, then (joint, camera-pair) solves, then the
// ordered per-joint reduction (see lift_person_kernel). The 4x4 null-vector problem of each pair is solved in
// registers: fp32 one-sided Jacobi SVD of A + one fp64 Rayleigh-quotient iteration (see triangulate_pair).
#include "common.cuh"

namespace b200pose {

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
//
// The right singular vector of the smallest singular value of the 4x4 DLT matrix A is found in two steps:
//   1. a register-resident one-sided (Hestenes) Jacobi SVD of A in fp32 - it works on A directly (no A^T A),
//      so it resolves the wanted vector to ~eps32 * s1 / (s3 - s4) even for the narrow-baseline ARP stereo
//      pair, at fp32 SFU/FMA rates;
//   2. ONE Rayleigh-quotient iteration in fp64 on M = A^T A: mu = x^T M x, solve (M - mu I) y = x by Gaussian
//      elimination with partial pivoting (pivots floored at 2^-60 |M|, so an exactly singular shift - noise-free
//      rays - still yields the null direction). RQI converges cubically: an fp32-accurate start lands at fp64
//      accuracy, within ~1e-9 relative of the fp64 SVD in the worst conditioned cases (s3/s1 ~ 1e-4, where the
//      reference's own result moves by as much for a 1-ulp change of its inputs) and ~1e-14 otherwise.
// The all-fp64 Jacobi this replaces cost ~10 k fp64 instructions per solve (sqrt/div chains) and made the
// encoder FP64-pipe bound; this version needs ~350 fp64 instructions plus ~2.5 k fp32 ones.
__device__ __forceinline__ void jacobi_null_vector_f32(const double (&A)[4][4], double (&x)[4])
{
    float a[4][4], v[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[i][k] = (float)A[i][k]; v[i][k] = (i == k) ? 1.0f : 0.0f; }
#pragma unroll 1
    for (int sweep = 0; sweep < 10; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                float al = 0.f, be = 0.f, ga = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) { al = fmaf(a[i][p], a[i][p], al); be = fmaf(a[i][q], a[i][q], be); ga = fmaf(a[i][p], a[i][q], ga); }
                if (ga * ga > 1e-13f * al * be && ga != 0.f) {
                    rotated = true;
                    const float zeta = (be - al) / (2.0f * ga);
                    const float t = copysignf(1.0f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.0f)));
                    const float c = rsqrtf(fmaf(t, t, 1.0f)), sn = c * t;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float ap = a[i][p], aq = a[i][q];
                        a[i][p] = c * ap - sn * aq; a[i][q] = sn * ap + c * aq;
                        const float vp = v[i][p], vq = v[i][q];
                        v[i][p] = c * vp - sn * vq; v[i][q] = sn * vp + c * vq;
                    }
                }
            }
        }
        if (!rotated) break;
    }
    int best = 0; float bn = 3.0e38f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float n = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) n = fmaf(a[i][q], a[i][q], n);
        if (n < bn) { bn = n; best = q; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float xi = v[i][0];
#pragma unroll
        for (int q = 1; q < 4; ++q) if (q == best) xi = v[i][q];
        x[i] = (double)xi;
    }
}

__device__ void triangulate_pair(const double* __restrict__ P1, const double* __restrict__ P2,
                                 double x1, double y1, double x2, double y2, double (&X)[3])
{
    double A[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        A[0][k] = x1 * P1[8 + k] - P1[k];
        A[1][k] = y1 * P1[8 + k] - P1[4 + k];
        A[2][k] = x2 * P2[8 + k] - P2[k];
        A[3][k] = y2 * P2[8 + k] - P2[4 + k];
    }
    double x[4];
    jacobi_null_vector_f32(A, x);
    // M = A^T A (symmetric), Rayleigh quotient of the fp32 vector
    double M[4][4];
    double mmax = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) {
            double m = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r) m = fma(A[r][i], A[r][j], m);
            M[i][j] = m; M[j][i] = m;
            mmax = fmax(mmax, fabs(m));
        }
    double xx = 0.0, xMx = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double mi = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) mi = fma(M[i][j], x[j], mi);
        xMx = fma(x[i], mi, xMx);
        xx = fma(x[i], x[i], xx);
    }
    const double mu = xMx / xx;
    const double floor_p = mmax * 8.673617379884035e-19;            // 2^-60 |M|
    // solve (M - mu I) y = x: Gaussian elimination with partial pivoting, rows kept in registers
    double B[4][5];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) B[i][j] = M[i][j] - (i == j ? mu : 0.0);
        B[i][4] = x[i];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int r = k + 1; r < 4; ++r) {                           // bring the largest |B[r][k]| of rows k.. to row k
            const bool sw = fabs(B[r][k]) > fabs(B[k][k]);
#pragma unroll
            for (int c = k; c < 5; ++c) {
                const double u = B[k][c], w = B[r][c];
                B[k][c] = sw ? w : u; B[r][c] = sw ? u : w;
            }
        }
        if (fabs(B[k][k]) < floor_p) B[k][k] = floor_p;
        const double inv = 1.0 / B[k][k];
#pragma unroll
        for (int r = k + 1; r < 4; ++r) {
            const double f = B[r][k] * inv;
#pragma unroll
            for (int c = k + 1; c < 5; ++c) B[r][c] = fma(-f, B[k][c], B[r][c]);
        }
    }
    // back substitution with a running rescale (y can be ~1e18 |x| when the shift is exact)
    double y[4];
    y[3] = B[3][4] / B[3][3];
    y[2] = (B[2][4] - B[2][3] * y[3]) / B[2][2];
    y[1] = (B[1][4] - B[1][2] * y[2] - B[1][3] * y[3]) / B[1][1];
    y[0] = (B[0][4] - B[0][1] * y[1] - B[0][2] * y[2] - B[0][3] * y[3]) / B[0][0];
    const bool ok = isfinite(y[0]) && isfinite(y[1]) && isfinite(y[2]) && isfinite(y[3]) && y[3] != 0.0;
    const double w3 = ok ? y[3] : x[3];
    X[0] = (ok ? y[0] : x[0]) / w3; X[1] = (ok ? y[1] : x[1]) / w3; X[2] = (ok ? y[2] : x[2]) / w3;
}

// ---- per-person lifting kernel ---------------------------------------------------------------------
// One CTA per person. The work of a person is re-parallelised over its natural units instead of one lane
// per joint:
//   phase U: thread per (camera slot, joint)  - fp64 undistortion; the encoder also forms the 10 per-view
//            numbers of that joint here and writes them into the shared-memory row
//   phase S: thread per (joint, camera pair)  - one register-resident 4x4 DLT solve each, results to
//            shared memory [joint][pair][3] (joints are processed in batches sized to 48 KB)
//   phase R: thread per joint                 - ordered reduction over the pairs (plain mean for the encoder,
//            upper median + 5 cm filter + mean for the baseline): same order as the reference's
//            itertools.combinations loop, so the result does not depend on the thread mapping
//   phase W: the whole CTA writes the person's row with 16-byte stores (fp32 and/or bf16 planes)
constexpr int kLiftThreads = 96;  // 170-180 pair solves per joint batch: two rounds of 96 waste fewer lanes than 128 + 52 (measured -4 %)
constexpr int kPairBudgetBytes = 48 * 1024;

struct LiftSmem {           // byte offsets into dynamic shared memory
    int ux, uy, have, cam, pa, pb, X, row, total;
    int n_pairs, joint_batch;
};
__host__ __device__ inline LiftSmem lift_smem(int n_slots, int row_len) {
    LiftSmem L;
    L.n_pairs = n_slots * (n_slots - 1) / 2;
    int jb = L.n_pairs > 0 ? kPairBudgetBytes / (L.n_pairs * 24) : kJ;
    L.joint_batch = jb < 1 ? 1 : (jb > kJ ? kJ : jb);
    int o = 0;
    L.ux = o; o += n_slots * kJ * 8;
    L.uy = o; o += n_slots * kJ * 8;
    L.X = o; o += (L.n_pairs > 0 ? L.joint_batch * L.n_pairs * 24 : 8);
    L.have = o; o += kJ * 4;
    L.cam = o; o += n_slots * 4;
    L.pa = o; o += (L.n_pairs + 1) * 2;
    L.pb = o; o += (L.n_pairs + 1) * 2;
    o = (o + 15) & ~15;
    L.row = o; o += ((row_len + 3) & ~3) * 4;
    L.total = o;
    return L;
}

template <bool ENCODE>
__global__ void __launch_bounds__(kLiftThreads) lift_person_kernel(
    int n_persons, const int* __restrict__ person_sk, const double* __restrict__ sk_xy, const float* __restrict__ sk_vp,
    const uint32_t* __restrict__ sk_mask, LiftTables t, int n_slots, int median_axis,
    float* __restrict__ x_f32, int ld_f32, __nv_bfloat16* __restrict__ x_hi, __nv_bfloat16* __restrict__ x_lo, int ld_planes,
    uint8_t* __restrict__ valid, double* __restrict__ xyz, uint8_t* __restrict__ mask, const int* __restrict__ n_dev)
{
    extern __shared__ __align__(16) unsigned char lift_raw[];
    __shared__ float red[kLiftThreads / 32];
    const int person = blockIdx.x;
    if (n_dev && person >= __ldg(n_dev)) return;            // launched for a capacity: the real count lives on the device
    const int tid = threadIdx.x;
    const int C = t.n_cameras;
    const int row_len = ENCODE ? kJ * 14 * n_slots : 0;
    const LiftSmem L = lift_smem(n_slots, row_len);
    double* ux = reinterpret_cast<double*>(lift_raw + L.ux);
    double* uy = reinterpret_cast<double*>(lift_raw + L.uy);
    double* X = reinterpret_cast<double*>(lift_raw + L.X);
    uint32_t* have = reinterpret_cast<uint32_t*>(lift_raw + L.have);
    int* cam_of_slot = reinterpret_cast<int*>(lift_raw + L.cam);
    unsigned short* pa = reinterpret_cast<unsigned short*>(lift_raw + L.pa);
    unsigned short* pb = reinterpret_cast<unsigned short*>(lift_raw + L.pb);
    float* row = reinterpret_cast<float*>(lift_raw + L.row);
    const int NP = L.n_pairs;

    // ---- setup: slot -> camera, pair table (lexicographic s1 < s2), zeroed row ----
    if (tid < n_slots) {
        int c = tid;                                     // baseline: slot = camera index (metrics_from_triangulation.py:239)
        if (ENCODE) {                                    // encoder: slots are used_cameras order (metrics_from_model.py:248-252)
            c = -1;
            for (int cc = 0; cc < C; ++cc) if (t.pe_slot[cc] == tid) c = cc;
        }
        cam_of_slot[tid] = c;
    }
    for (int i = tid; i < NP; i += kLiftThreads) {
        int a = 0, rem = i;                              // pair i -> (a, b)
        while (rem >= n_slots - 1 - a) { rem -= n_slots - 1 - a; ++a; }
        pa[i] = (unsigned short)a; pb[i] = (unsigned short)(a + 1 + rem);
    }
    if (tid < kJ) have[tid] = 0;
    if (ENCODE) for (int i = tid; i < ((row_len + 3) & ~3); i += kLiftThreads) row[i] = 0.f;
    __syncthreads();

    // ---- phase U: undistortion (+ the per-view numbers of the encoder) ----
    for (int i = tid; i < n_slots * kJ; i += kLiftThreads) {
        const int s = i / kJ, j = i - s * kJ;
        const int c = cam_of_slot[s];
        const int sk = (c >= 0) ? person_sk[(size_t)person * C + c] : -1;
        if (sk < 0 || !((sk_mask[sk] >> j) & 1u)) continue;
        const double x = sk_xy[(size_t)sk * 36 + 2 * j], y = sk_xy[(size_t)sk * 36 + 2 * j + 1];
        double nx, ny;
        undistort(x, y, t.k64 + 4 * c, t.dist64 + 5 * c, nx, ny);
        ux[i] = nx; uy[i] = ny;
        atomicOr(&have[j], 1u << s);
        if (ENCODE) {
            float* o = row + (size_t)s * (kJ * 14) + (size_t)j * 14;
            const double w2 = (double)t.W / 2.0, h2 = (double)t.Hh / 2.0;
            o[0] = sk_vp[(size_t)sk * 36 + 2 * j];
            o[1] = __double2float_rn((x - w2) / w2);
            o[2] = __double2float_rn((y - h2) / h2);
            o[3] = sk_vp[(size_t)sk * 36 + 2 * j + 1];
            const float* T = t.t_cam2root32 + 16 * c;
            const float fx_ = __double2float_rn(nx), fy_ = __double2float_rn(ny);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                o[4 + r] = __fdiv_rn(T[4 * r + 3], 10.0f);
                float acc = __fmul_rn(T[4 * r], fx_);
                acc = __fmaf_rn(T[4 * r + 1], fy_, acc);
                acc = __fmaf_rn(T[4 * r + 2], 1.0f, acc);
                acc = __fmaf_rn(T[4 * r + 3], 0.0f, acc);
                o[7 + r] = __fdiv_rn(acc, 10.0f);
            }
        }
    }
    __syncthreads();

    // ---- phases S + R, joints in batches ----
    const int j_first = ENCODE ? 1 : 0;                  // the encoder's hint skips joint 0 (dataset.py:75: `pos[0] > 0.`)
    for (int jb0 = j_first; jb0 < kJ; jb0 += L.joint_batch) {
        const int jb1 = min(kJ, jb0 + L.joint_batch);
        for (int i = tid; i < (jb1 - jb0) * NP; i += kLiftThreads) {
            const int jl = i / NP, pi = i - jl * NP;
            const int j = jb0 + jl;
            const int s1 = pa[pi], s2 = pb[pi];
            const uint32_t hv = have[j];
            if (!((hv >> s1) & 1u) || !((hv >> s2) & 1u)) continue;
            double Xp[3];
            triangulate_pair(t.p64 + 12 * cam_of_slot[s1], t.p64 + 12 * cam_of_slot[s2],
                             ux[s1 * kJ + j], uy[s1 * kJ + j], ux[s2 * kJ + j], uy[s2 * kJ + j], Xp);
            double* o = X + (size_t)i * 3;
            o[0] = Xp[0]; o[1] = Xp[1]; o[2] = Xp[2];
        }
        __syncthreads();
        if (tid < jb1 - jb0) {
            const int j = jb0 + tid;
            const uint32_t hv = have[j];
            const double* Xj = X + (size_t)tid * NP * 3;
            if (ENCODE) {
                if (__popc(hv) >= 2) {
                    double acc[3] = {0, 0, 0};
                    int n = 0;
                    for (int pi = 0; pi < NP; ++pi) {
                        if (!((hv >> pa[pi]) & 1u) || !((hv >> pb[pi]) & 1u)) continue;
                        acc[0] += Xj[3 * pi]; acc[1] += Xj[3 * pi + 1]; acc[2] += Xj[3 * pi + 2];
                        ++n;
                    }
                    float tri[4];
                    tri[0] = 1.0f;
#pragma unroll
                    for (int r = 0; r < 3; ++r) tri[1 + r] = __double2float_rn((acc[r] / (double)n) / 10.0);
                    for (int s = 0; s < n_slots; ++s) {  // every used-camera block (dataset.py:280-285)
                        float* o = row + (size_t)s * (kJ * 14) + (size_t)j * 14 + 10;
                        o[0] = tri[0]; o[1] = tri[1]; o[2] = tri[2]; o[3] = tri[3];
                    }
                }
            } else {
                double* out = xyz + ((size_t)person * kJ + j) * 3;
                if (__popc(hv) < 2) {
                    out[0] = out[1] = out[2] = 0.0;
                    mask[(size_t)person * kJ + j] = 0;
                } else {
                    // upper median of the median-axis coordinate by counting (ties by pair order), utils.py:70-71
                    int npairs = 0;
                    for (int pi = 0; pi < NP; ++pi) npairs += ((hv >> pa[pi]) & 1u) && ((hv >> pb[pi]) & 1u);
                    const int target = npairs / 2;
                    double med = 0.0;
                    for (int a = 0; a < NP; ++a) {
                        if (!((hv >> pa[a]) & 1u) || !((hv >> pb[a]) & 1u)) continue;
                        const double va = Xj[3 * a + median_axis];
                        int less = 0, eq_before = 0;
                        for (int k = 0; k < NP; ++k) {
                            if (!((hv >> pa[k]) & 1u) || !((hv >> pb[k]) & 1u)) continue;
                            const double vk = Xj[3 * k + median_axis];
                            less += vk < va;
                            eq_before += (vk == va) && (k < a);
                        }
                        if (less + eq_before == target) med = va;
                    }
                    double acc[3] = {0, 0, 0};
                    int kept = 0;
                    for (int pi = 0; pi < NP; ++pi) {
                        if (!((hv >> pa[pi]) & 1u) || !((hv >> pb[pi]) & 1u)) continue;
                        if (fabs(Xj[3 * pi + median_axis] - med) < 0.05) {
                            acc[0] += Xj[3 * pi]; acc[1] += Xj[3 * pi + 1]; acc[2] += Xj[3 * pi + 2];
                            ++kept;
                        }
                    }
                    out[0] = acc[0] / kept; out[1] = acc[1] / kept; out[2] = acc[2] / kept;
                    mask[(size_t)person * kJ + j] = 1;
                }
            }
        }
        __syncthreads();
    }
    if (!ENCODE) return;

    // ---- phase W: the person's row, 16-byte stores ----
    float abs_sum = 0.f;
    if (x_hi) {
        for (int v8 = tid; v8 < ld_planes / 8; v8 += kLiftThreads) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c0 = 8 * v8 + 2 * u;
                const float a = c0 < row_len ? row[c0] : 0.f, b = c0 + 1 < row_len ? row[c0 + 1] : 0.f;
                split_pack2(a, b, h[u], l[u]);
            }
            *reinterpret_cast<uint4*>(x_hi + (size_t)person * ld_planes + 8 * v8) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(x_lo + (size_t)person * ld_planes + 8 * v8) = make_uint4(l[0], l[1], l[2], l[3]);
        }
    }
    const bool vec_ok = x_f32 && (ld_f32 % 4 == 0) && ((reinterpret_cast<uintptr_t>(x_f32) & 15) == 0) && (row_len % 4 == 0);
    for (int v4 = tid; v4 < (row_len + 3) / 4; v4 += kLiftThreads) {
        const float4 v = *reinterpret_cast<const float4*>(row + 4 * v4);
        abs_sum += fabsf(v.x) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w);      // row is zero padded to a multiple of 4
        if (vec_ok) *reinterpret_cast<float4*>(x_f32 + (size_t)person * ld_f32 + 4 * v4) = v;
        else if (x_f32) {
            const float e[4] = {v.x, v.y, v.z, v.w};
            for (int u = 0; u < 4; ++u) if (4 * v4 + u < row_len) x_f32[(size_t)person * ld_f32 + 4 * v4 + u] = e[u];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) abs_sum += __shfl_xor_sync(0xffffffffu, abs_sum, o);
    if ((tid & 31) == 0) red[tid >> 5] = abs_sum;
    __syncthreads();
    if (valid && tid == 0) {
        float sum = 0.f;
        for (int i = 0; i < kLiftThreads / 32; ++i) sum += red[i];
        valid[person] = sum > 1.0f ? 1 : 0;                                  // dataset.py:287
    }
}

}  // namespace b200pose

using namespace b200pose;

static LiftTables make_tables(const b200pose_cameras* cams) {
    LiftTables t;
    t.n_cameras = cams->n_cameras; t.v_pe = cams->v_pe; t.W = cams->image_width; t.Hh = cams->image_height;
    t.pe_slot = cams->pe_slot; t.t_cam2root32 = cams->t_cam2root32; t.k64 = cams->k64; t.dist64 = cams->dist64; t.p64 = cams->p64;
    return t;
}

static int encode_persons_impl(int32_t n_persons, const int32_t* n_persons_dev, const int32_t* person_sk, const double* sk_xy, const float* sk_vp,
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
    const LiftSmem L = lift_smem(cams->v_pe, row_len);
    B2_CHECK_CUDA(cudaFuncSetAttribute(lift_person_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    lift_person_kernel<true><<<n_persons, kLiftThreads, L.total, (cudaStream_t)stream>>>(
        n_persons, person_sk, sk_xy, sk_vp, sk_mask, make_tables(cams), cams->v_pe, 0, x_f32, ld_f32,
        reinterpret_cast<__nv_bfloat16*>(x_hi), reinterpret_cast<__nv_bfloat16*>(x_lo), ld_planes, valid, nullptr, nullptr, n_persons_dev);
    B2_CHECK_LAUNCH();
    return B200POSE_OK;
}

extern "C" __attribute__((visibility("default"))) int b200pose_encode_persons(int32_t n_persons, const int32_t* person_sk, const double* sk_xy, const float* sk_vp,
                                       const uint32_t* sk_mask, const b200pose_cameras* cams,
                                       float* x_f32, int32_t ld_f32, uint16_t* x_hi, uint16_t* x_lo, int32_t ld_planes,
                                       uint8_t* valid, void* stream)
{
    return encode_persons_impl(n_persons, nullptr, person_sk, sk_xy, sk_vp, sk_mask, cams, x_f32, ld_f32, x_hi, x_lo, ld_planes, valid, stream);
}

extern "C" __attribute__((visibility("default"))) int b200pose_encode_persons_n(int32_t capacity, const int32_t* n_persons_dev, const int32_t* person_sk,
                                         const double* sk_xy, const float* sk_vp, const uint32_t* sk_mask, const b200pose_cameras* cams,
                                         float* x_f32, int32_t ld_f32, uint16_t* x_hi, uint16_t* x_lo, int32_t ld_planes,
                                         uint8_t* valid, void* stream)
{
    B2_CHECK_ARG(n_persons_dev, "encode_persons_n: null count pointer");
    return encode_persons_impl(capacity, n_persons_dev, person_sk, sk_xy, sk_vp, sk_mask, cams, x_f32, ld_f32, x_hi, x_lo, ld_planes, valid, stream);
}

extern "C" __attribute__((visibility("default"))) int b200pose_triangulate(int32_t n_persons, const int32_t* person_sk, const double* sk_xy,
                                    const uint32_t* sk_mask, const b200pose_cameras* cams, int32_t median_axis,
                                    double* xyz, uint8_t* mask, void* stream)
{
    B2_CHECK_ARG(person_sk && sk_xy && sk_mask && cams && xyz && mask, "triangulate: null pointer");
    B2_CHECK_ARG(median_axis >= 0 && median_axis < 3, "triangulate: median_axis must be 0..2");
    B2_CHECK_ARG(cams->n_cameras <= B200POSE_MAX_CAMERAS, "triangulate: too many cameras");
    if (n_persons == 0) return B200POSE_OK;
    const LiftSmem L = lift_smem(cams->n_cameras, 0);
    B2_CHECK_CUDA(cudaFuncSetAttribute(lift_person_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    lift_person_kernel<false><<<n_persons, kLiftThreads, L.total, (cudaStream_t)stream>>>(
        n_persons, person_sk, sk_xy, nullptr, sk_mask, make_tables(cams), cams->n_cameras, median_axis, nullptr, 0,
        nullptr, nullptr, 0, nullptr, xyz, mask, nullptr);
    B2_CHECK_LAUNCH();
    return B200POSE_OK;
}
