// io_kernels.cuh -- the data-parallel steps either side of the path: edge discretisation before it,
// structured (SoA) output and viewer buffers after it.
#pragma once
#include "kernels.cuh"

namespace tmesh {

// ---------------------------------------------------------------------------------------------------
// Structured output (the step right after the path, src/core/cgns.zig:69-101, 110-161): the AoS block (x,y interleaved,
// j fastest) as two SoA arrays with i fastest -- what cg_coord_write / cg_field_write take.  A 32x32 tile transpose
// through shared memory: coalesced 16 B loads along j, coalesced 8 B stores along i.  HBM-bound, 32 B per node.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) aos_to_soa_kernel(int ni, int nj, const double2* __restrict__ in, double* __restrict__ x, double* __restrict__ y) {
    __shared__ double2 tile[32][33];
    const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, j = j0 + tx;
        if (i < ni && j < nj) tile[r][tx] = in[(size_t)i * nj + j];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = j0 + r, i = i0 + tx;
        if (i < ni && j < nj) {
            const double2 v = tile[tx][r];
            x[(size_t)j * ni + i] = v.x;
            y[(size_t)j * ni + i] = v.y;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Edge.combine and projectNormal, the two other edge operations of the O4H blocking, batched: one
// CTA per job.
//   combine         discrete.zig:38-91 with the views of :94-136: the views' points back to back without the duplicated
//                   joints; clustering = offset of the view + (c[i] - c[first]) -- for a reversed view the differences
//                   still run from min(start, end) upwards (reference behaviour, :119-135) -- normalised by the last value
//   projectNormal   templates/O4H.zig:531-574: x_i + d n, n = (t_y, -t_x) / |t|, t central (one-sided at the two ends)
// Round-to-nearest intrinsics in the reference's operation order: bit-exact.
// ---------------------------------------------------------------------------------------------------
struct CombineView { int64_t src_off; int32_t start, end; };     // src_off: first point of the source edge in the input arrays
struct CombineJob { int64_t out_off; int32_t view_begin, n_views, n, _pad; };
__global__ void __launch_bounds__(128) edge_combine_kernel(const CombineJob* __restrict__ jobs, const CombineView* __restrict__ views,
                                                            const double2* __restrict__ src_pts, const double* __restrict__ src_cl,
                                                            double2* __restrict__ out_pts, double* __restrict__ out_cl) {
    const CombineJob job = jobs[blockIdx.x];
    __shared__ double s_init[16];   // clustering value at the first node of every view
    __shared__ int s_first[17];     // first output index of every view
    __shared__ double s_last;
    if (threadIdx.x == 0) {
        double value = 0.0;
        int pos = 0;
        for (int v = 0; v < job.n_views; ++v) {
            const CombineView w = views[job.view_begin + v];
            const int first = min(w.start, w.end), last = max(w.start, w.end);
            s_init[v] = value;
            s_first[v] = pos;
            value = __dadd_rn(value, __dsub_rn(src_cl[w.src_off + last], src_cl[w.src_off + first]));
            pos += last - first;       // the joint is shared with the next view
        }
        s_first[job.n_views] = pos;
        s_last = value;
    }
    __syncthreads();
    for (int v = 0; v < job.n_views; ++v) {
        const CombineView w = views[job.view_begin + v];
        const int first = min(w.start, w.end), len = abs(w.start - w.end) + 1;
        const int step = w.start > w.end ? -1 : 1;
        const double c_first = src_cl[w.src_off + first];
        for (int k = threadIdx.x; k < len; k += blockDim.x) {
            if (k == len - 1 && v + 1 < job.n_views) continue;  // the joint is written by the NEXT view (discrete.zig:66-70 overwrites it)
            out_pts[job.out_off + s_first[v] + k] = src_pts[w.src_off + w.start + step * k];
            const double u = k == 0 ? s_init[v] : __dadd_rn(s_init[v], __dsub_rn(src_cl[w.src_off + first + k], c_first));
            out_cl[job.out_off + s_first[v] + k] = __ddiv_rn(u, s_last);
        }
    }
}

struct ProjectJob { int64_t off; int32_t n, _pad; double distance; };
__global__ void __launch_bounds__(128) project_normal_kernel(const ProjectJob* __restrict__ jobs, const double2* __restrict__ in, double2* __restrict__ out) {
    const ProjectJob job = jobs[blockIdx.x];
    const double2* e = in + job.off;
    for (int i = threadIdx.x; i < job.n; i += blockDim.x) {
        double tx, ty;
        if (i == 0) { tx = __dsub_rn(e[1].x, e[0].x); ty = __dsub_rn(e[1].y, e[0].y); }
        else if (i == job.n - 1) { tx = __dsub_rn(e[i].x, e[i - 1].x); ty = __dsub_rn(e[i].y, e[i - 1].y); }
        else { tx = __dmul_rn(0.5, __dsub_rn(e[i + 1].x, e[i - 1].x)); ty = __dmul_rn(0.5, __dsub_rn(e[i + 1].y, e[i - 1].y)); }
        const double inv = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty))));
        const double nx = __dmul_rn(inv, ty), ny = __dmul_rn(inv, -tx);
        out[job.off + i] = make_double2(__dadd_rn(e[i].x, __dmul_rn(job.distance, nx)), __dadd_rn(e[i].y, __dmul_rn(job.distance, ny)));
    }
}

// ---------------------------------------------------------------------------------------------------
// Edge discretisation, the step right before the path: discrete.Edge.init =
// clustering.create + Curve.interpolate (src/core/discrete.zig:17-31), batched -- one CTA per edge, one thread per point.
//   clustering   clustering.zig:9-17 (uniform), :24-42 (Roberts), :56-95 (Vinokur tanh; delta comes from the host)
//   line         geometry.zig:26-40
//   spline       FittingSpline.interpolate = eval(paramAtArcFraction(u)), spline.zig:74-81, 112-139, 202-222, on the tables of
//                an already fitted spline (params, points, second derivatives, arc-length table)
// The arithmetic of the curves uses round-to-nearest intrinsics in the reference's operation order (no FMA contraction):
// with a uniform clustering the points are bit-exact; pow / tanh of the other clusterings are CUDA's libm (<= 2 ulp).
// ---------------------------------------------------------------------------------------------------
struct EdgeJob {
    int64_t out_off;        // first point of the edge in the output arrays
    int64_t spline_off;     // spline tables in `tables`: params[m], points[2m], zx[m], zy[m], arc[n_samples]
    int32_t n, curve, clustering, spline_m, n_samples, _pad;
    double line[4];         // start x,y ; end x,y
    double alpha, beta, delta, total_length;
};
__device__ __forceinline__ double edge_clustering(const EdgeJob& e, int i) {
    const double n_1 = (double)(e.n - 1);
    const double u = __ddiv_rn((double)i, n_1);
    if (e.clustering == 1) {  // Roberts
        const double tmp = pow(__ddiv_rn(__dadd_rn(e.beta, 1.0), __dsub_rn(e.beta, 1.0)), __ddiv_rn(__dsub_rn(u, e.alpha), __dsub_rn(1.0, e.alpha)));
        const double tbar = __dadd_rn(__dsub_rn(__dmul_rn(__dadd_rn(e.beta, __dmul_rn(2.0, e.alpha)), tmp), e.beta), __dmul_rn(2.0, e.alpha));
        return __ddiv_rn(tbar, __dmul_rn(__dadd_rn(__dmul_rn(2.0, e.alpha), 1.0), __dadd_rn(1.0, tmp)));
    }
    if (e.clustering == 2 && i > 0)  // Vinokur: 1 + tanh(delta/2 (u - 1)) / tanh(delta/2)
        return __dadd_rn(1.0, __ddiv_rn(tanh(__dmul_rn(__dmul_rn(0.5, e.delta), __dsub_rn(u, 1.0))), tanh(__dmul_rn(0.5, e.delta))));
    return u;
}
// FittingSpline.eval (spline.zig:202-222) on the tables of a fitted spline, reference operation order
__device__ __forceinline__ double2 spline_eval_param(const double* params, const double* points, const double* zx, const double* zy, int m, double param) {
    const double uu = param < 0.0 ? 0.0 : (param > 1.0 ? 1.0 : param);
    // idx = number of knots params[1..] below uu (the reference scans linearly), at most m-2
    int lo = 0, hi = m - 1;
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (params[mid + 1] < uu) lo = mid + 1; else hi = mid;
    }
    const int idx = lo >= m - 1 ? m - 2 : lo;
    const double h = __dsub_rn(params[idx + 1], params[idx]);
    const double a = __ddiv_rn(__dsub_rn(params[idx + 1], uu), h), b = __ddiv_rn(__dsub_rn(uu, params[idx]), h);
    const double a3 = __dsub_rn(__dmul_rn(__dmul_rn(a, a), a), a), b3 = __dsub_rn(__dmul_rn(__dmul_rn(b, b), b), b);
    const double hh = __dmul_rn(h, h);
    auto comp = [&](double y0, double y1, double z0, double z1) {
        const double lin = __dadd_rn(__dmul_rn(a, y0), __dmul_rn(b, y1));
        const double cub = __ddiv_rn(__dmul_rn(__dadd_rn(__dmul_rn(a3, z0), __dmul_rn(b3, z1)), hh), 6.0);
        return __dadd_rn(lin, cub);
    };
    return make_double2(comp(points[2 * idx], points[2 * idx + 2], zx[idx], zx[idx + 1]), comp(points[2 * idx + 1], points[2 * idx + 3], zy[idx], zy[idx + 1]));
}
__device__ __forceinline__ double2 edge_spline_point(const EdgeJob& e, const double* __restrict__ tables, double u) {
    const int m = e.spline_m;
    const double* params = tables + e.spline_off;
    const double* points = params + m;
    const double* zx = points + 2 * m;
    const double* zy = zx + m;
    const double* arc = zy + m;
    double param = 0.0;
    if (e.total_length != 0.0) {  // paramAtArcFraction
        const double target = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
        int lo = 0, hi = e.n_samples - 1;
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if (arc[mid] < target) lo = mid + 1; else hi = mid;
        }
        if (lo > 0) {
            const double a0 = arc[lo - 1], a1 = arc[lo];
            const double ns = (double)(e.n_samples - 1);
            const double p0 = __ddiv_rn((double)(lo - 1), ns), p1 = __ddiv_rn((double)lo, ns);
            const double t = a1 > a0 ? __ddiv_rn(__dsub_rn(target, a0), __dsub_rn(a1, a0)) : 0.0;
            param = __dadd_rn(p0, __dmul_rn(t, __dsub_rn(p1, p0)));
        }
    }
    return spline_eval_param(params, points, zx, zy, m, param);
}
__global__ void __launch_bounds__(128) edge_discretize_kernel(const EdgeJob* __restrict__ jobs, const double* __restrict__ tables, double2* __restrict__ points,
                                                              double* __restrict__ clustering) {
    const EdgeJob e = jobs[blockIdx.x];
    for (int i = threadIdx.x; i < e.n; i += blockDim.x) {
        const double u = edge_clustering(e, i);
        double2 p;
        if (e.curve == 0) {
            p.x = __dadd_rn(e.line[0], __dmul_rn(u, __dsub_rn(e.line[2], e.line[0])));
            p.y = __dadd_rn(e.line[1], __dmul_rn(u, __dsub_rn(e.line[3], e.line[1])));
        } else {
            p = edge_spline_point(e, tables, u);
        }
        clustering[e.out_off + i] = u;
        points[e.out_off + i] = p;
    }
}

// ---------------------------------------------------------------------------------------------------
// Spline fit, the first step of the blocking (spline.FittingSpline.init, src/core/spline.zig:24-110,
// 141-200): chord-length parameters, natural-cubic second derivatives by the tridiagonal recurrence, the 201-entry arc-length
// table.  One CTA per spline: the recurrences are serial (lanes 0 / 1 take x / y), the table's samples are evaluated by all
// threads and summed by one in the reference's left-to-right order -- bit-exact.
// ---------------------------------------------------------------------------------------------------
struct SplineFitJob { int64_t pt_off, arc_off; int32_t n, n_samples; };
__global__ void __launch_bounds__(128) spline_fit_kernel(const SplineFitJob* jobs, const double2* pts_all, double* params_all, double* zx_all, double* zy_all,
                                                          double* tmp_all /* 2 per point */, double* arc_all, double2* sample_all, double* total_length, int* err) {
    const SplineFitJob job = jobs[blockIdx.x];
    const double2* pts = pts_all + job.pt_off;
    double* params = params_all + job.pt_off;
    double* z[2] = {zx_all + job.pt_off, zy_all + job.pt_off};
    double* arc = arc_all + job.arc_off;
    double2* samples = sample_all + job.arc_off;
    const int n = job.n, m = job.n_samples;
    auto dist = [](double2 a, double2 b) {
        const double dx = __dsub_rn(b.x, a.x), dy = __dsub_rn(b.y, a.y);
        return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    };
    if (threadIdx.x == 0) {  // computeChordParams, spline.zig:141-155
        double total = 0.0;
        params[0] = 0.0;
        for (int i = 1; i < n; ++i) { total = __dadd_rn(total, dist(pts[i - 1], pts[i])); params[i] = total; }
        for (int i = 0; i < n; ++i) params[i] = total == 0.0 ? __ddiv_rn((double)i, (double)(n - 1)) : __ddiv_rn(params[i], total);
    }
    __syncthreads();
    if (threadIdx.x < 2) {  // computeSecondDerivs, spline.zig:157-200
        const int c = threadIdx.x;
        double* zc = z[c];
        double* tmp = tmp_all + 2 * job.pt_off + (size_t)c * n;
        auto y = [&](int i) { return c == 0 ? pts[i].x : pts[i].y; };
        for (int i = 0; i < n; ++i) { zc[i] = 0.0; tmp[i] = 0.0; }
        if (n > 2) {
            for (int i = 1; i < n - 1; ++i) {
                const double h_im1 = __dsub_rn(params[i], params[i - 1]), h_i = __dsub_rn(params[i + 1], params[i]);
                if (h_im1 == 0.0 || h_i == 0.0) { *err = 1; break; }   // CoincidentParameters
                const double alpha = __dsub_rn(__ddiv_rn(__dsub_rn(y(i + 1), y(i)), h_i), __ddiv_rn(__dsub_rn(y(i), y(i - 1)), h_im1));
                const double denom = __dsub_rn(__dmul_rn(2.0, __dsub_rn(params[i + 1], params[i - 1])), __dmul_rn(h_im1, tmp[i - 1]));
                tmp[i] = __ddiv_rn(h_i, denom);
                zc[i] = __ddiv_rn(__dsub_rn(__dmul_rn(6.0, alpha), __dmul_rn(h_im1, zc[i - 1])), denom);
            }
            zc[n - 1] = 0.0;
            for (int k = n - 2; k >= 0; --k) zc[k] = __dsub_rn(zc[k], __dmul_rn(tmp[k], zc[k + 1]));
        }
    }
    __syncthreads();
    const double* flat = reinterpret_cast<const double*>(pts);
    for (int k = threadIdx.x; k < m; k += blockDim.x)  // the samples of the arc-length table, spline.zig:87-110
        samples[k] = spline_eval_param(params, flat, z[0], z[1], n, __ddiv_rn((double)k, (double)(m - 1)));
    __syncthreads();
    if (threadIdx.x == 0) {
        double length = 0.0;
        arc[0] = 0.0;
        for (int k = 1; k < m; ++k) { length = __dadd_rn(length, dist(samples[k - 1], samples[k])); arc[k] = length; }
        total_length[blockIdx.x] = length;
        for (int k = 0; k < m; ++k) arc[k] = length == 0.0 ? 0.0 : __ddiv_rn(arc[k], length);
    }
}

// ---------------------------------------------------------------------------------------------------
// Viewer buffers (src/gui/lib.zig:227-318): f32 copies of all points with their x / y ranges
// (createPointBuffer) and the wireframe line indices (createWireframeElementBuffer: per block first the segments along j,
// then those along i), produced on the device -- into a mapped GL buffer if the caller hands one in.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) viewer_points_kernel(int64_t n, const double2* __restrict__ x, float2* __restrict__ out, float* __restrict__ partials /* grid x 4 */) {
    float xmin = 3.402823466e+38f, xmax = 1.175494351e-38f, ymin = 3.402823466e+38f, ymax = 1.175494351e-38f;  // floatMax / floatMin as in the reference
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (int64_t)gridDim.x * 256) {
        const double2 p = x[k];
        const float2 f = make_float2(__double2float_rn(p.x), __double2float_rn(p.y));
        out[k] = f;
        xmin = fminf(xmin, f.x); xmax = fmaxf(xmax, f.x); ymin = fminf(ymin, f.y); ymax = fmaxf(ymax, f.y);
    }
    __shared__ float sh[4][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh[0][w] = xmin; sh[1][w] = xmax; sh[2][w] = ymin; sh[3][w] = ymax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) { xmin = fminf(xmin, sh[0][q]); xmax = fmaxf(xmax, sh[1][q]); ymin = fminf(ymin, sh[2][q]); ymax = fmaxf(ymax, sh[3][q]); }
        float* p = partials + (size_t)blockIdx.x * 4;
        p[0] = xmin; p[1] = xmax; p[2] = ymin; p[3] = ymax;
    }
}
__global__ void viewer_ranges_kernel(const float* __restrict__ partials, int n_part, float* __restrict__ ranges /* xmin xmax ymin ymax */) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float r[4] = {3.402823466e+38f, 1.175494351e-38f, 3.402823466e+38f, 1.175494351e-38f};
    for (int k = 0; k < n_part; ++k) {
        r[0] = fminf(r[0], partials[4 * k]); r[1] = fmaxf(r[1], partials[4 * k + 1]);
        r[2] = fminf(r[2], partials[4 * k + 2]); r[3] = fmaxf(r[3], partials[4 * k + 3]);
    }
    for (int k = 0; k < 4; ++k) ranges[k] = r[k];
}
struct ViewerBlock { int64_t point_off, index_off; int32_t ni, nj; };  // offsets of the block in the point / index buffers
__global__ void __launch_bounds__(256) viewer_wireframe_kernel(const ViewerBlock* __restrict__ blocks, uint2* __restrict__ lines) {
    const ViewerBlock b = blocks[blockIdx.y];
    const int64_t n_j = (int64_t)b.ni * (b.nj - 1), n_i = (int64_t)b.nj * (b.ni - 1);   // segments along j, along i
    uint2* out = lines + b.index_off / 2;
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n_j + n_i; k += (int64_t)gridDim.x * 256) {
        unsigned p, q;
        if (k < n_j) {              // i-th line, j-th segment: (i*nj + j, i*nj + j + 1)
            const int64_t i = k / (b.nj - 1), j = k - i * (b.nj - 1);
            p = (unsigned)(b.point_off + i * b.nj + j); q = p + 1u;
        } else {                    // j-th column, i-th segment: (i*nj + j, (i+1)*nj + j)
            const int64_t kk = k - n_j, j = kk / (b.ni - 1), i = kk - j * (b.ni - 1);
            p = (unsigned)(b.point_off + i * b.nj + j); q = p + (unsigned)b.nj;
        }
        out[k] = make_uint2(p, q);
    }
}

// tm_mesh_download_block_async: device memory -> page-locked host memory mapped into the device's address space (16-byte stores
// over the host link; a grid-stride loop of a few CTAs is enough to fill it)
__global__ void __launch_bounds__(256) host_store_kernel(double2* __restrict__ dst, const double2* __restrict__ src, int64_t n) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) dst[k] = __ldcs(src + k);
}

}  // namespace tmesh
