// cb2_cells.cu — builds the support cells of cb2_cells.cuh on the device at cb2_shapes_upload.
// One thread per (polyhedron, cell); f64 arithmetic so the exclusion test itself adds no
// meaningful rounding (the margin is 4e-6 relative, f64 error is 1e-16).
#include <vector>

#include "cb2_cells.cuh"
#include "cb2_gjk.cuh"
#include <cstdio>
#include <cstdlib>

#include "cb2_kernels.h"

namespace cb2 {

struct d3 {
    double x, y, z;
};
__device__ __forceinline__ double ddot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// direction of cube-map coordinates (face, u, v): major component +-1, minors (u, v) in the
// cyclic order cell_index() uses
__device__ __forceinline__ d3 cube_dir(uint32_t face, double u, double v) {
    const uint32_t axis = face >> 1;
    const double s = (face & 1u) ? -1.0 : 1.0;
    d3 d;
    if (axis == 0) { d.x = s; d.y = u; d.z = v; }
    else if (axis == 1) { d.y = s; d.z = u; d.x = v; }
    else { d.z = s; d.x = u; d.y = v; }
    return d;
}

constexpr int CELL_SAMPLES = 9;  // 3 x 3 grid of dominator candidates per cell

__global__ void __launch_bounds__(128) k_build_cells(const dshape* __restrict__ shapes,
                                                    const float4* __restrict__ verts,
                                                    uint32_t n_shapes, uint4* __restrict__ cells,
                                                    uint16_t* __restrict__ ext, uint32_t ext_cap,
                                                    uint32_t* __restrict__ ext_cursor) {
    const uint32_t si = blockIdx.x;
    if (si >= n_shapes) return;
    const dshape sh = shapes[si];
    if (sh.kind != K_POLY) return;
    const uint32_t N = sh.cell_cfg & 0xFFu;
    const bool wide = (sh.cell_cfg >> 8) & 1u;
    if (N == 0) return;
    const uint32_t cell = blockIdx.y * blockDim.x + threadIdx.x;
    if (cell >= 6u * N * N) return;
    const uint32_t face = cell / (N * N), iu = (cell / N) % N, iv = cell % N;
    const float4* __restrict__ P = verts + sh.voff;
    const uint32_t V = sh.vcnt;
    const double R = (double)sh.hx;

    const double u0 = -1.0 + 2.0 * iu / N - CB2_CELL_PAD, u1 = -1.0 + 2.0 * (iu + 1) / N + CB2_CELL_PAD;
    const double v0 = -1.0 + 2.0 * iv / N - CB2_CELL_PAD, v1 = -1.0 + 2.0 * (iv + 1) / N + CB2_CELL_PAD;
    d3 smp[CELL_SAMPLES];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
            smp[a * 3 + b] = cube_dir(face, a == 0 ? u0 : (a == 1 ? 0.5 * (u0 + u1) : u1),
                                      b == 0 ? v0 : (b == 1 ? 0.5 * (v0 + v1) : v1));
    // corners are samples 0, 2, 6, 8
    const int corner_of[4] = {0, 2, 6, 8};

    // pass 1: support vertex of every sample direction = dominator candidates
    double best[CELL_SAMPLES];
    uint32_t wi[CELL_SAMPLES];
#pragma unroll
    for (int k = 0; k < CELL_SAMPLES; ++k) {
        best[k] = -1.0e300;
        wi[k] = 0;
    }
    for (uint32_t j = 0; j < V; ++j) {
        const float4 p = __ldg(P + j);
        const d3 q = {(double)p.x, (double)p.y, (double)p.z};
#pragma unroll
        for (int k = 0; k < CELL_SAMPLES; ++k) {
            const double t = ddot(q, smp[k]);
            if (t > best[k]) {
                best[k] = t;
                wi[k] = j;
            }
        }
    }
    // dominators' dots at the four corners, and the margin there
    double wc[CELL_SAMPLES][4], thr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const d3 c = smp[corner_of[i]];
        thr[i] = CB2_CELL_MARGIN * R * sqrt(ddot(c, c));
    }
#pragma unroll
    for (int k = 0; k < CELL_SAMPLES; ++k) {
        const float4 p = __ldg(P + wi[k]);
        const d3 q = {(double)p.x, (double)p.y, (double)p.z};
#pragma unroll
        for (int i = 0; i < 4; ++i) wc[k][i] = ddot(q, smp[corner_of[i]]);
    }

    // pass 2: keep every vertex no single dominator beats at all four corners
    const uint32_t inline_cap = wide ? CELL_INLINE16 : CELL_INLINE8;
    const uint32_t bits = wide ? 16u : 8u;
    uint32_t rec[4] = {0u, 0u, 0u, 0u};
    uint32_t cnt = 0;
    for (int pass = 0; pass < 2; ++pass) {
        uint32_t ext_off = 0;
        if (pass == 1) {
            // list does not fit the record: move it to the extension pool
            ext_off = atomicAdd(ext_cursor, cnt);
            if (ext_off + cnt > ext_cap) {  // pool exhausted: this cell falls back to the full scan
                rec[0] = wide ? 0xFFFEu : CELL_FULL;
                rec[1] = rec[2] = rec[3] = 0u;
                break;
            }
            rec[0] = wide ? 0xFFFFu : CELL_EXT;
            rec[1] = ext_off;
            rec[2] = cnt;
            rec[3] = 0u;
        }
        uint32_t n = 0;
        for (uint32_t j = 0; j < V; ++j) {
            const float4 p = __ldg(P + j);
            const d3 q = {(double)p.x, (double)p.y, (double)p.z};
            double pc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pc[i] = ddot(q, smp[corner_of[i]]);
            bool excluded = false;
#pragma unroll
            for (int k = 0; k < CELL_SAMPLES; ++k) {
                bool all = true;
#pragma unroll
                for (int i = 0; i < 4; ++i) all = all && (wc[k][i] - pc[i] > thr[i]);
                excluded = excluded || all;
            }
            // a NaN/Inf vertex fails every comparison and is kept (R is then non-finite too and
            // the lookup takes the full scan)
            if (!excluded) {
                if (pass == 0) {
                    if (n < inline_cap) {
                        const uint32_t slot = n + 1u;  // slot 0 holds the count
                        const uint32_t bitpos = slot * bits;
                        rec[bitpos >> 5] |= j << (bitpos & 31u);
                    }
                } else {
                    ext[ext_off + n] = (uint16_t)j;
                }
                ++n;
            }
        }
        if (pass == 0) {
            cnt = n;
            if (cnt <= inline_cap) {
                rec[0] |= cnt;
                break;
            }
        }
    }
    cells[sh.cell_off + cell] = make_uint4(rec[0], rec[1], rec[2], rec[3]);
}

// choose the cube-map resolution of a polyhedron with V vertices
uint32_t cells_resolution(uint32_t V) {
    if (V <= 8 || V > 65535) return 0;  // tiny: the scan is the list; huge: u16 indices do not fit
    // CB2_CELL_N="a,b,c": resolutions for V <= 64, V <= 512 and beyond (experiments)
    static const int* ovr = [] {
        static int v[3] = {0, 0, 0};
        const char* e = std::getenv("CB2_CELL_N");
        if (e && std::sscanf(e, "%d,%d,%d", &v[0], &v[1], &v[2]) == 3 && v[0] > 0 && v[1] > 0 && v[2] > 0 &&
            v[0] < 64 && v[1] < 64 && v[2] < 64)
            return (const int*)v;
        return (const int*)nullptr;
    }();
    if (ovr) return (uint32_t)(V <= 64 ? ovr[0] : (V <= 512 ? ovr[1] : ovr[2]));
    if (V <= 64) return 4;
    if (V <= 512) return 8;
    return 16;
}

cudaError_t build_support_cells(const dshape* d_shapes, const float4* d_verts, uint32_t n_shapes,
                                uint32_t max_N, uint4* d_cells, uint16_t* d_ext, uint32_t ext_cap,
                                uint32_t* d_ext_cursor, cudaStream_t st) {
    if (n_shapes == 0 || max_N == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_ext_cursor, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    const uint32_t max_cells = 6u * max_N * max_N;
    // gridDim.x is limited to 2^31-1, gridDim.y to 65535: shapes on x
    dim3 grid(n_shapes, (max_cells + 127u) / 128u);
    k_build_cells<<<grid, 128, 0, st>>>(d_shapes, d_verts, n_shapes, d_cells, d_ext, ext_cap,
                                        d_ext_cursor);
    return cudaGetLastError();
}

}  // namespace cb2
