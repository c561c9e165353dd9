// Stage 1: skeleton-candidate graph (COO + CSR) and node features, batched over frames.
//
// Reference behaviour restated (paths relative to the reference root):
//   skeleton_matching/graph_generator.py:573-605  load_people_view_graph  (head order, nodes_camera)
//   skeleton_matching/graph_generator.py:627-656  add_edge_node_to_graph  (5 edges per edge-node)
//   skeleton_matching/graph_generator.py:813-876  process_test            (pair loop order)
//   skeleton_matching/graph_generator.py:444-508  initializeWithAlternative3 (feature row)
//
// The graph of a frame has a closed form given the sizes n_0..n_{G-1} of its camera groups (runs of
// equal camera among the heads, in frame-dict order): edge-node k of pair block (gi<gj) joins head
// h1 = start[gi] + r / n_gj and h2 = start[gj] + r % n_gj. One CTA per frame writes every array
// directly - no DGL graph object, no O(M^2 F) feature concatenation.
#include "common.cuh"

namespace b200pose {

constexpr int kMaxGroups = B200POSE_MAX_CAMERAS;
constexpr int kMaxBlocks = kMaxGroups * (kMaxGroups - 1) / 2;

struct FrameGroups {
    int G;
    int H;
    int M;
    int start[kMaxGroups];
    int size[kMaxGroups];
    int cam[kMaxGroups];
    int deg_prefix[kMaxGroups + 1];   // CSR offset of the first head of group g
    int blk_base[kMaxBlocks + 1];     // edge-node index of the first edge-node of pair block i
    short blk_gi[kMaxBlocks];
    short blk_gj[kMaxBlocks];
};

__device__ __forceinline__ int block_index(int gi, int gj, int G) {
    return gi * G - gi * (gi + 1) / 2 + (gj - gi - 1);
}

// thread 0 of the CTA fills the group tables of one frame (H is tens of heads)
__device__ void compute_groups(FrameGroups& fg, const int* __restrict__ sk_cam, int h0, int H) {
    int G = 0;
    for (int i = 0; i < H; ++i) {
        int c = sk_cam[h0 + i];
        if (G == 0 || c != fg.cam[G - 1]) {
            if (G < kMaxGroups) {
                fg.cam[G] = c;
                fg.start[G] = i;
                fg.size[G] = 0;
                ++G;
            }
        }
        fg.size[G - 1]++;
    }
    fg.G = G;
    fg.H = H;
    int nb = 0, base = 0;
    for (int gi = 0; gi < G; ++gi)
        for (int gj = gi + 1; gj < G; ++gj) {
            fg.blk_gi[nb] = (short)gi;
            fg.blk_gj[nb] = (short)gj;
            fg.blk_base[nb] = base;
            base += fg.size[gi] * fg.size[gj];
            ++nb;
        }
    fg.blk_base[nb] = base;
    fg.M = base;
    int acc = 0;
    for (int g = 0; g < G; ++g) {
        fg.deg_prefix[g] = acc;
        acc += fg.size[g] * (1 + H - fg.size[g]);
    }
    fg.deg_prefix[G] = acc;   // == H + 2M
}

__global__ void __launch_bounds__(128) build_graph_kernel(
    int n_frames, const int* __restrict__ head_off, const int* __restrict__ node_off,
    const int* __restrict__ sk_cam, const int* __restrict__ sm_slot,
    int* __restrict__ src, int* __restrict__ dst, int* __restrict__ row_ptr, int* __restrict__ col,
    int* __restrict__ pairs, int* __restrict__ node_cam)
{
    __shared__ FrameGroups fg;
    const int b = blockIdx.x;
    const int h0 = head_off[b];
    const int H = head_off[b + 1] - h0;
    const int n0 = node_off[b];
    const int Nb = node_off[b + 1] - n0;
    const int m0 = n0 - h0;                     // first edge-node (global edge-node index) of the frame
    const int e0 = h0 + 5 * m0;                 // first edge of the frame
    if (threadIdx.x == 0) compute_groups(fg, sk_cam, h0, H);
    __syncthreads();
    const int G = fg.G;
    const int M = min(fg.M, Nb - H);            // the packer computed node_off with the same formula
    const int nblk = G * (G - 1) / 2;

    // heads: self loops, node camera, CSR rows
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        if (src) { src[e0 + h] = h; dst[e0 + h] = h; }
        if (node_cam) node_cam[n0 + h] = sm_slot[sk_cam[h0 + h]];
        if (row_ptr || col) {
            int g = 0;
            while (g + 1 < G && h >= fg.start[g + 1]) ++g;
            const int i = h - fg.start[g];
            const int ng = fg.size[g];
            int pos = fg.deg_prefix[g] + i * (1 + H - ng);
            if (row_ptr) row_ptr[n0 + h] = e0 + pos;
            if (col) {
                col[e0 + pos++] = n0 + h;                                   // self loop (edge id h)
                for (int gi = 0; gi < g; ++gi) {                            // blocks (gi, g): h is head2
                    const int base = n0 + H + fg.blk_base[block_index(gi, g, G)];
                    for (int r = 0; r < fg.size[gi]; ++r) col[e0 + pos++] = base + r * ng + i;
                }
                for (int gj = g + 1; gj < G; ++gj) {                        // blocks (g, gj): h is head1
                    const int base = n0 + H + fg.blk_base[block_index(g, gj, G)] + i * fg.size[gj];
                    for (int r = 0; r < fg.size[gj]; ++r) col[e0 + pos++] = base + r;
                }
            }
        }
    }
    // edge-nodes
    const int csr_enode0 = fg.deg_prefix[G];
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        int lo = 0, hi = nblk - 1;                                           // last block with base <= k
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (fg.blk_base[mid] <= k) lo = mid; else hi = mid - 1;
        }
        const int gi = fg.blk_gi[lo], gj = fg.blk_gj[lo];
        const int r = k - fg.blk_base[lo];
        const int h1 = fg.start[gi] + r / fg.size[gj];
        const int h2 = fg.start[gj] + r % fg.size[gj];
        const int e = H + k;
        if (src) {
            int* s = src + e0 + H + 5 * k;
            int* d = dst + e0 + H + 5 * k;
            s[0] = h1; d[0] = e;      // graph_generator.py:632-635
            s[1] = e;  d[1] = h1;     // :636-639
            s[2] = h2; d[2] = e;      // :640-643
            s[3] = e;  d[3] = h2;     // :644-647
            s[4] = e;  d[4] = e;      // :648-651
        }
        if (pairs) { pairs[2 * (m0 + k)] = h1; pairs[2 * (m0 + k) + 1] = h2; }
        if (node_cam) node_cam[n0 + e] = -1;
        const int pos = csr_enode0 + 3 * k;
        if (row_ptr) row_ptr[n0 + e] = e0 + pos;
        if (col) {
            col[e0 + pos] = n0 + h1;
            col[e0 + pos + 1] = n0 + h2;
            col[e0 + pos + 2] = n0 + e;
        }
    }
    if (row_ptr && b == n_frames - 1 && threadIdx.x == 0) row_ptr[n0 + Nb] = e0 + H + 5 * M;
}

// ---------------------------------------------------------------------------------------------
// Graph from an explicit edge-node list: the training-side topology of process_training
// (graph_generator.py:672-810), where edge-nodes follow the people / other-people / spurious loops, both
// (h1, h2) and (h2, h1) exist, and dgl.batch concatenates several such graphs - no closed form, but the same
// wiring per edge-node (add_edge_node_to_graph, :627-656). One CTA per graph: in-degrees by shared-memory atomics
// (a count, so order-free), exclusive scan, then one thread per head walks the edge-node list in order, so the
// in-edges of every node are in ascending reference edge id like the closed-form builder's.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_graph_pairs_kernel(
    int n_graphs, int max_heads, const int* __restrict__ head_off, const int* __restrict__ node_off,
    const int* __restrict__ sk_cam, const int* __restrict__ sm_slot, const int* __restrict__ pairs,
    int* __restrict__ src, int* __restrict__ dst, int* __restrict__ row_ptr, int* __restrict__ col, int* __restrict__ node_cam)
{
    extern __shared__ int off_s[];                  // [H + 1] in-degree, then CSR offset of every head
    const int b = blockIdx.x;
    const int h0 = head_off[b], H = head_off[b + 1] - h0;
    const int n0 = node_off[b], Nb = node_off[b + 1] - n0;
    const int M = Nb - H, m0 = n0 - h0, e0 = h0 + 5 * m0;
    const int* pr = pairs + 2 * (size_t)m0;
    if (H > max_heads || M < 0) return;             // outside the sized plan (the host checked the batch-wide maximum)
    for (int h = threadIdx.x; h <= H; h += blockDim.x) off_s[h] = h < H ? 1 : 0;      // the self loop
    __syncthreads();
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        const unsigned h1 = (unsigned)pr[2 * k], h2 = (unsigned)pr[2 * k + 1];
        if (h1 < (unsigned)H) atomicAdd(&off_s[h1], 1);    // a pair that names a head outside the graph gets no in-edge
        if (h2 < (unsigned)H) atomicAdd(&off_s[h2], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int h = 0; h <= H; ++h) { const int d = off_s[h]; off_s[h] = run; run += d; }
    }
    __syncthreads();
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        int pos = e0 + off_s[h];
        if (row_ptr) row_ptr[n0 + h] = pos;
        if (node_cam) node_cam[n0 + h] = sm_slot[sk_cam[h0 + h]];
        if (src) { src[e0 + h] = h; dst[e0 + h] = h; }
        if (col) {
            col[pos++] = n0 + h;                                             // self loop (edge id h)
            for (int k = 0; k < M; ++k)                                       // (e -> h1) is edge H + 5k + 1, (e -> h2) edge H + 5k + 3
                if (pr[2 * k] == h || pr[2 * k + 1] == h) col[pos++] = n0 + H + k;
        }
    }
    const int csr_enode0 = off_s[H];                                         // == H + 2M
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        const int h1 = pr[2 * k], h2 = pr[2 * k + 1], e = H + k;
        if (src) {
            int* s = src + e0 + H + 5 * k;
            int* d = dst + e0 + H + 5 * k;
            s[0] = h1; d[0] = e; s[1] = e; d[1] = h1; s[2] = h2; d[2] = e; s[3] = e; d[3] = h2; s[4] = e; d[4] = e;   // :632-651
        }
        if (node_cam) node_cam[n0 + e] = -1;
        const int pos = e0 + csr_enode0 + 3 * k;
        if (row_ptr) row_ptr[n0 + e] = pos;
        if (col) { col[pos] = n0 + h1; col[pos + 1] = n0 + h2; col[pos + 2] = n0 + e; }
    }
    if (row_ptr && b == n_graphs - 1 && threadIdx.x == 0) row_ptr[n0 + Nb] = e0 + H + 5 * M;
}

// ---------------------------------------------------------------------------------------------
// Node features. feature_value() evaluates one column of a head row with exactly the reference's
// arithmetic: i/j normalisation in float64 then rounded to fp32 (graph_generator.py:496-497), rays as
// two tiny fp32 matmuls whose CPU summation order is a sequential FMA chain (:488-489).
// ---------------------------------------------------------------------------------------------
struct FeatureTables {
    const int* sm_slot;
    const float* kinv32;
    const float* t_cam2root32;
    float W, Hh;
    int F;
};

__device__ __forceinline__ float feature_value(const FeatureTables& t, int col, int cam, int slot, uint32_t mask,
                                               const double* __restrict__ xy, const float* __restrict__ vp)
{
    if (col == 0) return 1.0f;
    const int rel = col - 2 - 180 * slot;
    if (rel < 0 || rel >= 180) return 0.0f;
    const int j = rel / 10, k = rel - 10 * j;
    if (!((mask >> j) & 1u)) return 0.0f;
    const double x = xy[2 * j], y = xy[2 * j + 1];
    switch (k) {
        case 0: { const double w2 = (double)t.W / 2.0; return __double2float_rn((x - w2) / w2); }
        case 1: { const double h2 = (double)t.Hh / 2.0; return __double2float_rn((h2 - y) / h2); }
        case 2: return vp[2 * j];
        case 3: return vp[2 * j + 1];
        case 4: case 5: case 6: return t.t_cam2root32[cam * 16 + (k - 4) * 4 + 3];
        default: {
            const float xf = __double2float_rn(x), yf = __double2float_rn(y);
            const float* Ki = t.kinv32 + cam * 9;
            float rc[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float acc = __fmul_rn(Ki[3 * i], xf);
                acc = __fmaf_rn(Ki[3 * i + 1], yf, acc);
                acc = __fmaf_rn(Ki[3 * i + 2], 1.0f, acc);
                rc[i] = acc;
            }
            const float* T = t.t_cam2root32 + cam * 16 + (k - 7) * 4;
            float acc = __fmul_rn(T[0], rc[0]);
            acc = __fmaf_rn(T[1], rc[1], acc);
            acc = __fmaf_rn(T[2], rc[2], acc);
            acc = __fmaf_rn(T[3], 0.0f, acc);
            return acc;
        }
    }
}

// dense fp32 rows for every node: one warp per node row
__global__ void __launch_bounds__(256) node_features_f32_kernel(
    int n_frames, int n_nodes_total, const int* __restrict__ head_off, const int* __restrict__ node_off,
    const double* __restrict__ sk_xy, const float* __restrict__ sk_vp, const uint32_t* __restrict__ sk_mask,
    const int* __restrict__ sk_cam, FeatureTables t, float* __restrict__ out, int ld)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_nodes_total) return;
    // frame of this node: binary search in node_off
    int lo = 0, hi = n_frames - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (node_off[mid] <= warp) lo = mid; else hi = mid - 1;
    }
    const int local = warp - node_off[lo];
    const int H = head_off[lo + 1] - head_off[lo];
    float* row = out + (size_t)warp * ld;
    if (local >= H) {                                        // edge-node: one-hot column 1 (:630)
        for (int c = lane; c < t.F; c += 32) row[c] = (c == 1) ? 1.0f : 0.0f;
        return;
    }
    const int s = head_off[lo] + local;
    const int cam = sk_cam[s];
    const int slot = t.sm_slot[cam];
    const uint32_t mask = sk_mask[s];
    const double* xy = sk_xy + (size_t)s * 36;
    const float* vp = sk_vp + (size_t)s * 36;
    for (int c = lane; c < t.F; c += 32) row[c] = feature_value(t, c, cam, slot, mask, xy, vp);
}

// planes for the S head rows + one shared edge-node row (row S): one warp per row, 2 columns per lane
__global__ void __launch_bounds__(256) head_features_planes_kernel(
    int n_heads_total, const double* __restrict__ sk_xy, const float* __restrict__ sk_vp,
    const uint32_t* __restrict__ sk_mask, const int* __restrict__ sk_cam, FeatureTables t,
    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int ld)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp > n_heads_total) return;
    uint32_t* rhi = reinterpret_cast<uint32_t*>(hi + (size_t)warp * ld);
    uint32_t* rlo = reinterpret_cast<uint32_t*>(lo + (size_t)warp * ld);
    const bool enode = (warp == n_heads_total);
    int cam = 0, slot = 0; uint32_t mask = 0;
    const double* xy = nullptr; const float* vp = nullptr;
    if (!enode) {
        cam = sk_cam[warp]; slot = t.sm_slot[cam]; mask = sk_mask[warp];
        xy = sk_xy + (size_t)warp * 36; vp = sk_vp + (size_t)warp * 36;
    }
    for (int c2 = lane; c2 < ld / 2; c2 += 32) {
        float v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int c = 2 * c2 + u;
            if (c >= t.F) v[u] = 0.f;
            else if (enode) v[u] = (c == 1) ? 1.0f : 0.0f;
            else v[u] = feature_value(t, c, cam, slot, mask, xy, vp);
        }
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(v[0], h0, l0);
        split_bf16(v[1], h1, l1);
        rhi[c2] = pack_bf16x2(h0, h1);
        rlo[c2] = pack_bf16x2(l0, l1);
    }
}

// fp32 matrix -> planes
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, int rows, int cols, int ld_in,
                                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                           int ld)
{
    const size_t total = (size_t)rows * (ld / 2);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / (ld / 2));
        const int c = 2 * (int)(i % (ld / 2));
        const float a = (c < cols) ? x[(size_t)r * ld_in + c] : 0.f;
        const float b = (c + 1 < cols) ? x[(size_t)r * ld_in + c + 1] : 0.f;
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(a, h0, l0);
        split_bf16(b, h1, l1);
        reinterpret_cast<uint32_t*>(hi)[(size_t)r * (ld / 2) + c / 2] = pack_bf16x2(h0, h1);
        reinterpret_cast<uint32_t*>(lo)[(size_t)r * (ld / 2) + c / 2] = pack_bf16x2(l0, l1);
    }
}

}  // namespace b200pose

using namespace b200pose;

extern "C" __attribute__((visibility("default"))) int b200pose_build_graph(int32_t n_frames, const int32_t* head_off, const int32_t* node_off,
                                    const int32_t* sk_cam, const b200pose_cameras* cams,
                                    int32_t* src, int32_t* dst, int32_t* row_ptr, int32_t* col,
                                    int32_t* pairs, int32_t* node_cam, void* stream)
{
    B2_CHECK_ARG(n_frames >= 0 && head_off && node_off && sk_cam && cams, "build_graph: null input");
    B2_CHECK_ARG((src == nullptr) == (dst == nullptr), "build_graph: src and dst go together");
    B2_CHECK_ARG(cams->n_cameras <= B200POSE_MAX_CAMERAS, "build_graph: more than %d cameras", B200POSE_MAX_CAMERAS);
    if (n_frames == 0) return B200POSE_OK;
    build_graph_kernel<<<n_frames, 128, 0, (cudaStream_t)stream>>>(n_frames, head_off, node_off, sk_cam, cams->sm_slot,
                                                                    src, dst, row_ptr, col, pairs, node_cam);
    B2_CHECK_LAUNCH();
    return B200POSE_OK;
}

extern "C" __attribute__((visibility("default"))) int b200pose_build_graph_pairs(int32_t n_graphs, const int32_t* head_off, const int32_t* node_off,
                                          const int32_t* sk_cam, const b200pose_cameras* cams, const int32_t* pairs,
                                          int32_t max_heads_per_graph,
                                          int32_t* src, int32_t* dst, int32_t* row_ptr, int32_t* col, int32_t* node_cam, void* stream)
{
    B2_CHECK_ARG(n_graphs >= 0 && head_off && node_off && sk_cam && cams && pairs, "build_graph_pairs: null input");
    B2_CHECK_ARG((src == nullptr) == (dst == nullptr), "build_graph_pairs: src and dst go together");
    B2_CHECK_ARG(max_heads_per_graph >= 0 && max_heads_per_graph <= 40000, "build_graph_pairs: max_heads_per_graph out of range");
    if (n_graphs == 0) return B200POSE_OK;
    const size_t smem = (size_t)(max_heads_per_graph + 1) * sizeof(int);
    B2_CHECK_CUDA(cudaFuncSetAttribute(build_graph_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    build_graph_pairs_kernel<<<n_graphs, 256, smem, (cudaStream_t)stream>>>(n_graphs, max_heads_per_graph, head_off, node_off, sk_cam, cams->sm_slot, pairs,
                                                                            src, dst, row_ptr, col, node_cam);
    B2_CHECK_LAUNCH();
    return B200POSE_OK;
}

extern "C" __attribute__((visibility("default"))) int b200pose_node_features(int32_t n_frames, int32_t n_heads_total, int32_t n_nodes_total, const int32_t* head_off,
                                      const int32_t* node_off, const double* sk_xy, const float* sk_vp,
                                      const uint32_t* sk_mask, const int32_t* sk_cam, const b200pose_cameras* cams,
                                      float* feats_f32, int32_t ld_f32, uint16_t* head_hi, uint16_t* head_lo,
                                      int32_t ld_planes, void* stream)
{
    B2_CHECK_ARG(cams && head_off && node_off && sk_xy && sk_vp && sk_mask && sk_cam, "node_features: null input");
    FeatureTables t;
    t.sm_slot = cams->sm_slot; t.kinv32 = cams->kinv32; t.t_cam2root32 = cams->t_cam2root32_sm;
    t.W = cams->image_width; t.Hh = cams->image_height; t.F = 2 + 180 * cams->v_sm;
    cudaStream_t st = (cudaStream_t)stream;
    if (feats_f32 && n_frames > 0) {
        B2_CHECK_ARG(ld_f32 >= t.F, "node_features: ld_f32 < F");
        if (n_nodes_total > 0) {
            const int warps_per_block = 8;
            node_features_f32_kernel<<<ceil_div(n_nodes_total, warps_per_block), 256, 0, st>>>(
                n_frames, n_nodes_total, head_off, node_off, sk_xy, sk_vp, sk_mask, sk_cam, t, feats_f32, ld_f32);
            B2_CHECK_LAUNCH();
        }
    }
    if (head_hi || head_lo) {
        B2_CHECK_ARG(head_hi && head_lo, "node_features: planes go together");
        B2_CHECK_ARG(ld_planes % 64 == 0 && ld_planes >= t.F, "node_features: ld_planes must be a multiple of 64 >= F");
        const int rows = n_heads_total + 1;
        head_features_planes_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(
            n_heads_total, sk_xy, sk_vp, sk_mask, sk_cam, t,
            reinterpret_cast<__nv_bfloat16*>(head_hi), reinterpret_cast<__nv_bfloat16*>(head_lo), ld_planes);
        B2_CHECK_LAUNCH();
    }
    return B200POSE_OK;
}

extern "C" __attribute__((visibility("default"))) int b200pose_split_planes(const float* x, int32_t rows, int32_t cols, int32_t ld_in,
                                     uint16_t* hi, uint16_t* lo, int32_t ld_planes, void* stream)
{
    B2_CHECK_ARG(x && hi && lo, "split_planes: null pointer");
    B2_CHECK_ARG(ld_planes % 64 == 0 && ld_planes >= cols && ld_in >= cols, "split_planes: bad leading dimensions");
    if (rows == 0) return B200POSE_OK;
    const size_t total = (size_t)rows * (ld_planes / 2);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_planes_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld_in,
        reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), ld_planes);
    B2_CHECK_LAUNCH();
    return B200POSE_OK;
}
