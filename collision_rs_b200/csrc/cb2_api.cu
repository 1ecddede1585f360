// cb2_api.cu — the extern "C" boundary declared in include/collision_b200.h.
//
// One ctx per GPU.  Host-buffer entry points stream the batch through the device in chunks: one
// copy-in stream, three compute streams taking the chunks in turn, one copy-out stream, tied by
// per-buffer events (see host_pipeline below; truly asynchronous when the caller's buffers are
// pinned).  Device-buffer entry points only enqueue kernels on the caller's stream.  No CPU
// fallback anywhere.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "cb2_gjk.cuh"
#include "cb2_kernels.h"

using namespace cb2;

#define CB2_NS_MAX 8
#ifndef CB2_DUMP0_DIV  // tier 0 -> 1 dump room: one pair in CB2_DUMP0_DIV
#define CB2_DUMP0_DIV 4
#endif
struct cb2_ctx {
    int device = 0;
    int sm_count = 0;
    std::string err;
    uint64_t launches = 0;
    // shape table
    dshape* d_shapes = nullptr;
    float4* d_verts = nullptr;
    uint4* d_cells = nullptr;        // support cells of all polyhedra (cb2_cells.cuh)
    uint16_t* d_cell_ext = nullptr;  // their extension pool
    uint32_t* d_cell_cursor = nullptr;
    uint32_t* d_bad_index = nullptr;  // set by the kernels when a pair names a shape outside the table
    uint32_t* h_bad_index = nullptr;  // pinned mirror the host-buffer entry points read back
    uint32_t n_shapes = 0;
    uint64_t n_verts = 0;
    // 2-D shape table (cb2_shapes2_upload), independent of the 3-D one
    cb2_shape2* d_shapes2 = nullptr;
    float* d_verts2 = nullptr;
    uint32_t n_shapes2 = 0;
    bool long_tails2 = false;  // see narrow2_args::long_tails (decided at cb2_shapes2_upload)
    // GJK2 -> EPA2 hand-off, one per host-API stream (device-API calls use slot 0)
    uint32_t* d_cnt2[CB2_NS_MAX] = {};
    uint32_t* d_hits2[CB2_NS_MAX] = {};
    void* d_sx2[CB2_NS_MAX] = {};
    uint64_t ws2_cap[CB2_NS_MAX] = {};
    // host-API staging
#ifndef CB2_NS
#define CB2_NS 8
#endif
    static constexpr int NS = CB2_NS;  // host-API pipeline depth (streams / staging buffers)
    cudaStream_t stream[NS] = {};
    cudaEvent_t done[NS] = {};
    // host-API pipeline events per staging buffer: inputs landed | kernels finished | buffer free
    cudaEvent_t ev_in[NS] = {}, ev_k[NS] = {}, ev_free[NS] = {};
    cudaEvent_t ev_tail[NS] = {};  // tier 0 finished -> the tail stream may start the later tiers
    // GJK-ahead pipeline: GJK of a chunk done | the chunk's workspace slot free again
    cudaEvent_t ev_front[NS] = {}, ev_wsfree[NS] = {};
    cudaStream_t s_front = nullptr;  // high-priority stream of the GJK stages
    cudaStream_t s_tail[2] = {nullptr, nullptr};  // EPA tiers 1, 2 and the second chance (above tier 0's priority)
    int tail_prio = 0;               // CB2_TAIL_PRIO=1: the later tiers on their own streams, above tier 0's
                                     // priority (measured: 23.1 against 21.0 ms per 4 M-pair call — their
                                     // latency-bound tails then hold SM resources tier 0 would use)
    int ahead = 1;                   // CB2_AHEAD=0: every stage of a chunk on the chunk's stream
    cudaStream_t s_in = nullptr, s_out = nullptr;  // the pipeline's copy-in / copy-out streams
    int n_cstreams = 3;  // compute streams the pipeline rotates over (CB2_NC overrides, <= NS)
    int ramp = 3;       // short chunks at the front of a batch: chunk/2^ramp ... chunk/2 (CB2_RAMP=h[,t]; 0 disables)
    int ramp_tail = 3;  // ... and at its end
    char* d_stage[NS] = {};
    size_t stage_cap = 0;
    sap_workspace* sap = nullptr;
    // composite table (cb2_composites_upload) and the composite-call workspace
    uint32_t* d_comp_off = nullptr;
    uint32_t* d_comp_shape = nullptr;
    float4* d_comp_local = nullptr;
    uint32_t n_comp = 0;
    std::vector<uint32_t> h_comp_off;
    // ... and the 2-D composite table (cb2_composites2_upload)
    uint32_t* d_comp2_off = nullptr;
    uint32_t* d_comp2_shape = nullptr;
    float4* d_comp2_local = nullptr;
    uint32_t n_comp2 = 0;
    std::vector<uint32_t> h_comp2_off;
    void* d_cx = nullptr;
    size_t cx_cap = 0;
    cb2_aabb3* d_lift = nullptr;  // cb2_sap2_find_collider_pairs_dev: the boxes lifted to 3-D
    uint64_t lift_cap = 0;
    uint32_t* d_body_shape = nullptr;  // body table of cb2_gjk3_intersection_bodies (host buffers)
    float4* d_body_t = nullptr;
    uint64_t body_cap = 0;
    void* d_misc = nullptr;  // host-API SAP / AABB buffers
    size_t misc_cap = 0;
    // classified (polyhedron) narrow phase workspaces: [0..NS) follow the host-API streams,
    // device-API calls use ws[0] (one submitting stream at a time per ctx)
    narrow_ws ws[NS] = {};
    uint64_t ws_cap[NS] = {};
    retry_ws retry[NS] = {};
    uint32_t* d_queue[NS] = {};  // work-queue cursors of the lane-refill kernels
    // optional stage timing (cb2_set_stage_timing): events around GJK | EPA tier 0 | EPA tier 1 |
    // second-chance, recorded on the launching stream of the LAST device-buffer intersection
    bool stage_timing = false;
    cudaEvent_t stage_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t stage_stream = nullptr;
    bool stage_valid = false;
    uint64_t chunk_pairs = 3u << 18;  // host-API pipeline granularity (CB2_CHUNK_PAIRS overrides)
    bool has_poly = false;
    bool legacy_epa = false;  // CB2_LEGACY_EPA=1: thread-per-pair EPA in local memory (A/B testing)
    bool no_cells = false;    // CB2_NO_CELLS=1: polyhedron support by full vertex scan (A/B testing)
};

static std::string g_create_err;

// CB2_TRACE=1: stage marks inside a host-pipeline chunk (events on the chunk's compute stream)
struct trace_mark {
    cudaEvent_t ev;
    const char* what;
};
static std::vector<trace_mark> g_trace;
static bool trace_on() {
    static const bool on = std::getenv("CB2_TRACE") && std::getenv("CB2_TRACE")[0] == '1';
    return on;
}
static void trace_stage(const char* what, cudaStream_t st) {
    if (!trace_on()) return;
    trace_mark m;
    cudaEventCreate(&m.ev);
    cudaEventRecord(m.ev, st);
    m.what = what;
    g_trace.push_back(m);
}

static void free_ws(narrow_ws& w) {  // the per-pair buffers
    for (int k = 0; k < EPA_TIERS; ++k) {
        if (w.list[k]) cudaFree(w.list[k]);
        w.list[k] = nullptr;
    }
    for (int k = 0; k + 1 < EPA_TIERS; ++k) {
        if (w.dump[k]) cudaFree(w.dump[k]);
        w.dump[k] = nullptr;
        w.dump_cap[k] = 0;
    }
    if (w.simplex) cudaFree(w.simplex);
    w.simplex = nullptr;
}

#define CK(ctx, x)                                                                     \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            (ctx)->err = std::string(#x) + ": " + cudaGetErrorString(e_);              \
            return CB2_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

static int fail(cb2_ctx* c, int code, const char* msg) {
    if (c) c->err = msg;
    return code;
}

extern "C" {

uint32_t cb2_abi_version(void) { return CB2_ABI_VERSION; }

void cb2_default_settings(cb2_settings* out) {
    out->distance_tolerance = 0.000001f;    // GJK_DISTANCE_TOLERANCE, gjk/mod.rs:23
    out->continuous_tolerance = 0.000001f;  // GJK_CONTINUOUS_TOLERANCE, gjk/mod.rs:24
    out->epa_tolerance = 0.00001f;          // EPA_TOLERANCE, epa/mod.rs:15
    out->max_iterations = 100;              // MAX_ITERATIONS, gjk/mod.rs:22, epa/mod.rs:16
}

int cb2_create(int device, cb2_ctx** out) {
    if (!out) return CB2_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        g_create_err = std::string("no usable CUDA device (") +
                       (e != cudaSuccess ? cudaGetErrorString(e) : "index out of range") +
                       "); this library has no CPU fallback";
        return CB2_ERR_NO_DEVICE;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        g_create_err = "cudaSetDevice failed";
        return CB2_ERR_NO_DEVICE;
    }
    cb2_ctx* c = new cb2_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        g_create_err = "cudaGetDeviceProperties failed";
        delete c;
        return CB2_ERR_CUDA;
    }
    c->sm_count = prop.multiProcessorCount;
    for (int i = 0; i < cb2_ctx::NS; ++i) {
        if (cudaStreamCreateWithFlags(&c->stream[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming) != cudaSuccess) {
            g_create_err = "stream/event creation failed";
            delete c;
            return CB2_ERR_CUDA;
        }
    }
    {
        const char* e1 = std::getenv("CB2_LEGACY_EPA");
        const char* e2 = std::getenv("CB2_NO_CELLS");
        const char* e3 = std::getenv("CB2_CHUNK_PAIRS");
        if (e3 && std::atoll(e3) > 0) c->chunk_pairs = (uint64_t)std::atoll(e3);
        const char* e4 = std::getenv("CB2_NC");
        if (e4 && std::atoi(e4) >= 1 && std::atoi(e4) <= cb2_ctx::NS) c->n_cstreams = std::atoi(e4);
        const char* e5 = std::getenv("CB2_RAMP");
        if (e5 && e5[0] >= '0' && e5[0] <= '5') {
            c->ramp = c->ramp_tail = e5[0] - '0';
            if (e5[1] == ',' && e5[2] >= '0' && e5[2] <= '5') c->ramp_tail = e5[2] - '0';
        }
        const char* e6 = std::getenv("CB2_AHEAD");
        if (e6) c->ahead = std::atoi(e6);
        const char* e7 = std::getenv("CB2_TAIL_PRIO");
        if (e7) c->tail_prio = std::atoi(e7);
        c->legacy_epa = e1 && e1[0] == '1';
        c->no_cells = e2 && e2[0] == '1';
    }
    c->sap = sap_workspace_create();
    if (!c->sap) {
        g_create_err = "pinned allocation failed";
        delete c;
        return CB2_ERR_CUDA;
    }
    *out = c;
    return CB2_OK;
}

void cb2_destroy(cb2_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int k = 0; k < 5; ++k)
        if (c->stage_ev[k]) cudaEventDestroy(c->stage_ev[k]);
    for (int i = 0; i < cb2_ctx::NS; ++i) {
        if (c->d_stage[i]) cudaFree(c->d_stage[i]);
        if (c->done[i]) cudaEventDestroy(c->done[i]);
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
        if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]);
        if (c->ev_tail[i]) cudaEventDestroy(c->ev_tail[i]);
        if (c->ev_front[i]) cudaEventDestroy(c->ev_front[i]);
        if (c->ev_wsfree[i]) cudaEventDestroy(c->ev_wsfree[i]);
        if (i == 0 && c->s_front) cudaStreamDestroy(c->s_front);
        if (i < 2 && c->s_tail[i]) cudaStreamDestroy(c->s_tail[i]);
        if (i == 0 && c->s_in) cudaStreamDestroy(c->s_in);
        if (i == 0 && c->s_out) cudaStreamDestroy(c->s_out);
        if (c->stream[i]) cudaStreamDestroy(c->stream[i]);
    }
    if (c->d_shapes) cudaFree(c->d_shapes);
    if (c->d_verts) cudaFree(c->d_verts);
    if (c->d_shapes2) cudaFree(c->d_shapes2);
    if (c->d_verts2) cudaFree(c->d_verts2);
    for (int i = 0; i < CB2_NS_MAX; ++i) {
        if (c->d_cnt2[i]) cudaFree(c->d_cnt2[i]);
        if (c->d_hits2[i]) cudaFree(c->d_hits2[i]);
        if (c->d_sx2[i]) cudaFree(c->d_sx2[i]);
    }
    if (c->d_cells) cudaFree(c->d_cells);
    if (c->d_cell_ext) cudaFree(c->d_cell_ext);
    if (c->d_cell_cursor) cudaFree(c->d_cell_cursor);
    if (c->d_bad_index) cudaFree(c->d_bad_index);
    if (c->h_bad_index) cudaFreeHost(c->h_bad_index);
    if (c->d_misc) cudaFree(c->d_misc);
    if (c->d_body_shape) cudaFree(c->d_body_shape);
    if (c->d_body_t) cudaFree(c->d_body_t);
    if (c->d_lift) cudaFree(c->d_lift);
    if (c->d_comp_off) cudaFree(c->d_comp_off);
    if (c->d_comp_shape) cudaFree(c->d_comp_shape);
    if (c->d_comp_local) cudaFree(c->d_comp_local);
    if (c->d_comp2_off) cudaFree(c->d_comp2_off);
    if (c->d_comp2_shape) cudaFree(c->d_comp2_shape);
    if (c->d_comp2_local) cudaFree(c->d_comp2_local);
    if (c->d_cx) cudaFree(c->d_cx);
    for (int i = 0; i < cb2_ctx::NS; ++i) {
        free_ws(c->ws[i]);
        if (c->d_queue[i]) cudaFree(c->d_queue[i]);
        if (c->ws[i].EC) cudaFree(c->ws[i].EC);
        if (c->ws[i].pa) cudaFree(c->ws[i].pa);
        if (c->retry[i].list) cudaFree(c->retry[i].list);
        if (c->retry[i].count) cudaFree(c->retry[i].count);
        if (c->retry[i].scratch) cudaFree(c->retry[i].scratch);
    }
    sap_workspace_destroy(c->sap);
    delete c;
}

const char* cb2_last_error(const cb2_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }
int cb2_device_sm_count(const cb2_ctx* c) { return c ? c->sm_count : 0; }
uint64_t cb2_kernel_launches(const cb2_ctx* c) { return c ? c->launches : 0; }

int cb2_set_stage_timing(cb2_ctx* c, int on) {
    if (!c) return CB2_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    if (on)
        for (int k = 0; k < 5; ++k)
            if (!c->stage_ev[k]) CK(c, cudaEventCreate(&c->stage_ev[k]));
    c->stage_timing = on != 0;
    c->stage_valid = false;
    return CB2_OK;
}

int cb2_last_stage_ms(cb2_ctx* c, float* out4) {
    if (!c || !out4) return CB2_ERR_ARG;
    if (!c->stage_valid) return fail(c, CB2_ERR_ARG, "no timed FullResolution intersection yet");
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaEventSynchronize(c->stage_ev[4]));
    for (int k = 0; k < 4; ++k) CK(c, cudaEventElapsedTime(&out4[k], c->stage_ev[k], c->stage_ev[k + 1]));
    return CB2_OK;
}

}  // extern "C"

// ConvexPolyhedron::new_with_faces: build_half_edges (polyhedron.rs:282-380) on the host, reduced
// to what hill_climb_support_point reads: per vertex, the targets of its ring walk in walk order.
// Fills rings[voff + v] (and ring_cnt) for the vertices of shape s.
static const char* build_rings(const cb2_shape& s, const float* verts_xyz, const uint32_t* faces,
                               uint64_t n_faces, std::vector<uint32_t>& ring_cnt,
                               std::vector<std::vector<uint32_t>>& rings) {
    uint32_t f_off, f_cnt;
    std::memcpy(&f_off, &s.p[0], 4);
    std::memcpy(&f_cnt, &s.p[1], 4);
    const uint32_t V = s.vtx_cnt;
    if (V == 0 || f_cnt == 0) return "HalfEdge polyhedron without vertices or faces";
    if (!faces || (uint64_t)f_off + f_cnt > n_faces) return "face range outside the face table";
    std::vector<uint32_t> vedge(V, 0), target, twin, next;
    std::vector<char> vready(V, 0);
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> edge_map;
    const float* vb = verts_xyz + 3ull * s.vtx_off;
    for (uint32_t f = 0; f < f_cnt; ++f) {
        const uint32_t* fv = faces + 3ull * (f_off + f);
        for (int k = 0; k < 3; ++k)
            if (fv[k] >= V) return "face vertex index out of range";
        {   // Plane::from_points(a, b, c).unwrap() (plane.rs:78-95): cross ~ 0 panics.  f32, unfused.
            volatile float e[6];
            for (int k = 0; k < 3; ++k) {
                e[k] = vb[3 * fv[1] + k] - vb[3 * fv[0] + k];
                e[3 + k] = vb[3 * fv[2] + k] - vb[3 * fv[0] + k];
            }
            volatile float m0 = e[1] * e[5], m1 = e[2] * e[4], m2 = e[2] * e[3], m3 = e[0] * e[5],
                           m4 = e[0] * e[4], m5 = e[1] * e[3];
            const float nx = m0 - m1, ny = m2 - m3, nz = m4 - m5;
            const float eps = 1.1920929e-7f;
            if (std::fabs(nx) <= eps && std::fabs(ny) <= eps && std::fabs(nz) <= eps)
                return "degenerate face: ConvexPolyhedron::new_with_faces panics (plane.rs:78-95)";
        }
        uint32_t face_edge[3] = {0, 0, 0};
        for (int j = 0; j < 3; ++j) {
            const int i = j == 0 ? 2 : j - 1;
            const uint32_t v0 = fv[i], v1 = fv[j];
            uint32_t e01, e10;
            auto it = edge_map.find({v0, v1});
            if (it != edge_map.end()) {
                e01 = it->second;
                e10 = twin[e01];
            } else {
                e01 = (uint32_t)target.size();
                e10 = e01 + 1;
                target.push_back(v1); twin.push_back(e10); next.push_back(0);
                target.push_back(v0); twin.push_back(e01); next.push_back(0);
            }
            edge_map[{v0, v1}] = e01;
            edge_map[{v1, v0}] = e10;
            if (!vready[v0]) {
                vedge[v0] = e01;
                vready[v0] = 1;
            }
            face_edge[i] = e01;
        }
        for (int j = 0; j < 3; ++j) next[face_edge[j == 0 ? 2 : j - 1]] = face_edge[j];
    }
    const size_t E = target.size();
    for (uint32_t v = 0; v < V; ++v) {
        std::vector<uint32_t> ring;
        uint32_t e = vedge[v];
        do {
            ring.push_back(target[e]);
            e = next[twin[e]];
            if (ring.size() > E) return "a vertex ring does not close (not a closed 2-manifold)";
        } while (e != vedge[v]);
        ring_cnt[s.vtx_off + v] = (uint32_t)ring.size();
        rings[s.vtx_off + v] = std::move(ring);
    }
    return nullptr;
}

// A composite table indexes the shape table it was uploaded against: a new shape table drops it
// (the caller uploads composites again), so a stale part index can never reach a kernel.
static void drop_composites(cb2_ctx* c, bool d2) {
    uint32_t*& off = d2 ? c->d_comp2_off : c->d_comp_off;
    uint32_t*& shp = d2 ? c->d_comp2_shape : c->d_comp_shape;
    float4*& loc = d2 ? c->d_comp2_local : c->d_comp_local;
    if (!off && !shp && !loc) return;
    cudaDeviceSynchronize();
    if (off) cudaFree(off);
    if (shp) cudaFree(shp);
    if (loc) cudaFree(loc);
    off = shp = nullptr;
    loc = nullptr;
    (d2 ? c->n_comp2 : c->n_comp) = 0;
    (d2 ? c->h_comp2_off : c->h_comp_off).clear();
}

static int shapes_upload(cb2_ctx* c, const cb2_shape* shapes, uint32_t n_shapes,
                         const float* verts_xyz, uint64_t n_verts, const uint32_t* faces,
                         uint64_t n_faces) {
    if (!c || (!shapes && n_shapes) || (!verts_xyz && n_verts)) return fail(c, CB2_ERR_ARG, "null argument");
    CK(c, cudaSetDevice(c->device));
    drop_composites(c, false);
    std::vector<dshape> hs(n_shapes);
    std::vector<uint32_t> ring_cnt;                // HalfEdge rings, per global vertex
    std::vector<std::vector<uint32_t>> rings;
    uint64_t total_cells = 0;
    uint32_t max_N = 0;
    c->has_poly = false;
    for (uint32_t i = 0; i < n_shapes; ++i) {
        const cb2_shape& s = shapes[i];
        dshape d;
        std::memset(&d, 0, sizeof(d));
        d.kind = s.kind;
        if (s.kind == CB2_SPHERE) {
            d.hx = s.p[0];
        } else if (s.kind == CB2_CUBOID) {
            // half_dim = dim / (1 + 1), cuboid.rs:36 (division by two is exact)
            d.hx = s.p[0] / 2.0f;
            d.hy = s.p[1] / 2.0f;
            d.hz = s.p[2] / 2.0f;
        } else if (s.kind == CB2_CUBE) {  // Cuboid::new(dim, dim, dim), cuboid.rs:151-157
            d.kind = CB2_CUBOID;
            d.hx = d.hy = d.hz = s.p[0] / 2.0f;
        } else if (s.kind == CB2_QUAD) {  // half_dim = dim / 2, quad.rs:36-42
            d.hx = s.p[0] / 2.0f;
            d.hy = s.p[1] / 2.0f;
        } else if (s.kind == CB2_CYLINDER || s.kind == CB2_CAPSULE) {
            d.hx = s.p[0];  // half_height
            d.hy = s.p[1];  // radius
        } else if (s.kind == CB2_PARTICLE) {
            // no parameters
        } else if (s.kind == CB2_POLYHEDRON_HE) {
            if ((uint64_t)s.vtx_off + s.vtx_cnt > n_verts)
                return fail(c, CB2_ERR_ARG, "polyhedron vertex range outside the vertex table");
            if (ring_cnt.empty()) {
                ring_cnt.assign(n_verts + 1, 0);
                rings.resize(n_verts);
            }
            const char* why = build_rings(s, verts_xyz, faces, n_faces, ring_cnt, rings);
            if (why) return fail(c, CB2_ERR_ARG, why);
            d.voff = s.vtx_off;
            d.vcnt = s.vtx_cnt;
        } else if (s.kind == CB2_POLYHEDRON) {
            if ((uint64_t)s.vtx_off + s.vtx_cnt > n_verts)
                return fail(c, CB2_ERR_ARG, "polyhedron vertex range outside the vertex table");
            d.voff = s.vtx_off;
            d.vcnt = s.vtx_cnt;
            c->has_poly = true;
            // support cells (cb2_cells.cuh): R = max |vertex| rounded up, cube-map resolution N
            double r2 = 0.0;
            for (uint32_t k = 0; k < s.vtx_cnt; ++k) {
                const float* v = verts_xyz + 3 * ((uint64_t)s.vtx_off + k);
                const double q = (double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2];
                if (!(q <= r2)) r2 = q;  // NaN sticks: the lookup then always takes the full scan
            }
            d.hx = std::nextafter((float)std::sqrt(r2), INFINITY);
            const uint32_t N = c->no_cells ? 0u : cells_resolution(s.vtx_cnt);
            if (N && total_cells + 6ull * N * N > 0xFFFFFFFFull)
                return fail(c, CB2_ERR_ARG, "support-cell table exceeds 2^32 records");
            d.cell_off = (uint32_t)total_cells;
            d.cell_cfg = N | ((s.vtx_cnt > 256 ? 1u : 0u) << 8);
            total_cells += 6ull * N * N;
            if (N > max_N) max_N = N;
        } else {
            return fail(c, CB2_ERR_ARG, "unknown shape kind");
        }
        hs[i] = d;
    }
    // vertex table: float4 (x, y, z, w) with w = offset of the vertex's HalfEdge ring record (in
    // uint32 units from the owning shape's first vertex); ring records [count, targets...] live
    // behind the vertices in the same buffer (cb2_cells.cuh: poly_hill_climb)
    std::vector<float4> hv(n_verts);
    for (uint64_t i = 0; i < n_verts; ++i)
        hv[i] = make_float4(verts_xyz[3 * i], verts_xyz[3 * i + 1], verts_xyz[3 * i + 2], 0.0f);
    std::vector<uint32_t> ring_data;
    if (!rings.empty()) {
        for (uint32_t i = 0; i < n_shapes; ++i) {
            if (shapes[i].kind != CB2_POLYHEDRON_HE) continue;
            const uint64_t v0 = shapes[i].vtx_off;
            for (uint32_t k = 0; k < shapes[i].vtx_cnt; ++k) {
                // uint32 index of this record, counted from the shape's first vertex
                const uint64_t rel = (n_verts - v0) * 4 + ring_data.size();
                if (rel > 0xFFFFFFFFull) return fail(c, CB2_ERR_ARG, "HalfEdge ring table exceeds 2^32 words");
                const uint32_t rel32 = (uint32_t)rel;
                std::memcpy(&hv[v0 + k].w, &rel32, 4);
                const std::vector<uint32_t>& ring = rings[v0 + k];
                ring_data.push_back((uint32_t)ring.size());
                ring_data.insert(ring_data.end(), ring.begin(), ring.end());
            }
        }
    }
    CK(c, cudaDeviceSynchronize());
    if (c->d_shapes) cudaFree(c->d_shapes);
    if (c->d_verts) cudaFree(c->d_verts);
    if (c->d_cells) cudaFree(c->d_cells);
    if (c->d_cell_ext) cudaFree(c->d_cell_ext);
    c->d_shapes = nullptr;
    c->d_verts = nullptr;
    c->d_cells = nullptr;
    c->d_cell_ext = nullptr;
    CK(c, cudaMalloc(&c->d_shapes, sizeof(dshape) * (n_shapes ? n_shapes : 1)));
    CK(c, cudaMalloc(&c->d_verts, sizeof(float4) * (n_verts ? n_verts : 1) + sizeof(uint32_t) * ring_data.size()));
    if (n_shapes)
        CK(c, cudaMemcpy(c->d_shapes, hs.data(), sizeof(dshape) * n_shapes, cudaMemcpyHostToDevice));
    if (n_verts)
        CK(c, cudaMemcpy(c->d_verts, hv.data(), sizeof(float4) * n_verts, cudaMemcpyHostToDevice));
    if (!ring_data.empty())
        CK(c, cudaMemcpy(reinterpret_cast<char*>(c->d_verts) + sizeof(float4) * n_verts, ring_data.data(),
                         sizeof(uint32_t) * ring_data.size(), cudaMemcpyHostToDevice));
    c->n_shapes = n_shapes;
    c->n_verts = n_verts;
    // build the support cells on the device (one thread per cell)
    const uint64_t ext_cap64 = total_cells * 4 + (1u << 16);
    const uint32_t ext_cap = ext_cap64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)ext_cap64;
    CK(c, cudaMalloc(&c->d_cells, sizeof(uint4) * (total_cells ? total_cells : 1)));
    CK(c, cudaMalloc(&c->d_cell_ext, sizeof(uint16_t) * (size_t)ext_cap));
    if (!c->d_cell_cursor) CK(c, cudaMalloc(&c->d_cell_cursor, sizeof(uint32_t)));
    if (total_cells) {
        cudaError_t e = build_support_cells(c->d_shapes, c->d_verts, n_shapes, max_N, c->d_cells,
                                            c->d_cell_ext, ext_cap, c->d_cell_cursor, nullptr);
        ++c->launches;
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            c->err = std::string("support-cell build: ") + cudaGetErrorString(e);
            return CB2_ERR_CUDA;
        }
    }
    return CB2_OK;
}

extern "C" {
int cb2_shapes_upload(cb2_ctx* c, const cb2_shape* shapes, uint32_t n_shapes,
                      const float* verts_xyz, uint64_t n_verts) {
    return shapes_upload(c, shapes, n_shapes, verts_xyz, n_verts, nullptr, 0);
}
int cb2_shapes_upload_faces(cb2_ctx* c, const cb2_shape* shapes, uint32_t n_shapes,
                            const float* verts_xyz, uint64_t n_verts, const uint32_t* faces,
                            uint64_t n_faces) {
    return shapes_upload(c, shapes, n_shapes, verts_xyz, n_verts, faces, n_faces);
}
}  // extern "C"

// ---- narrow phase ---------------------------------------------------------------------------------
enum op_kind { OP_INTERSECTION, OP_DISTANCE, OP_TOI, OP_INTERSECT, OP_MANIFOLD };

static int check_settings(cb2_ctx* c, const cb2_settings* s, bool dim2 = false) {
    if (!s) return fail(c, CB2_ERR_ARG, "settings is null");
    if (s->max_iterations > (dim2 ? CB2_MAX_ITERATIONS_2D : CB2_MAX_ITERATIONS))
        return fail(c, CB2_ERR_UNSUPPORTED,
                    dim2 ? "max_iterations above CB2_MAX_ITERATIONS_2D (100)" : "max_iterations above CB2_MAX_ITERATIONS (250)");
    return CB2_OK;
}

static int ensure_ws(cb2_ctx* c, int slot, uint64_t n) {
    narrow_ws& w = c->ws[slot];
    if (!w.EC) CK(c, cudaMalloc(&w.EC, sizeof(epa_counters)));
    if (!w.pa) CK(c, cudaMalloc(&w.pa, epa_pa_scratch_bytes(c->sm_count)));
    if (c->ws_cap[slot] < n) {
        free_ws(w);
        c->ws_cap[slot] = 0;
        for (int k = 0; k < EPA_TIERS; ++k) CK(c, cudaMalloc(&w.list[k], n * sizeof(uint32_t)));
        CK(c, cudaMalloc(&w.simplex, n * 8 * sizeof(float4)));
        // room for 1 pair in 4 (tier 0 -> 1) and 1 in 32 (tier 1 -> 2) to hand their polytope
        // over; beyond that a pair restarts in the next tier instead of resuming
        for (int k = 0; k + 1 < EPA_TIERS; ++k) {
            const uint64_t frac = k == 0 ? n / CB2_DUMP0_DIV : n / 32;
            const uint64_t dc = frac > 4096 ? frac : (n < 4096 ? n : 4096);
            CK(c, cudaMalloc(&w.dump[k], dc * epa_dump_record_bytes(k)));
            w.dump_cap[k] = (uint32_t)dc;
        }
        c->ws_cap[slot] = n;
    }
    return CB2_OK;
}

static int ensure_retry(cb2_ctx* c, int slot) {
    retry_ws& r = c->retry[slot];
    if (!r.list) CK(c, cudaMalloc(&r.list, sizeof(uint32_t) * RETRY_LIST_CAP));
    if (!r.count) CK(c, cudaMalloc(&r.count, 2 * sizeof(uint32_t)));
    if (!r.scratch) CK(c, cudaMalloc(&r.scratch, retry_scratch_bytes()));
    return CB2_OK;
}

// phase: 0 = the whole operation; the GJK-ahead host pipeline issues FullResolution intersections in
// two parts on two streams: 1 = counters + GJK stage, 2 = EPA tiers + second chance
static int launch_op_inner(cb2_ctx* c, op_kind op, uint32_t strategy, narrow_args& A,
                           cudaStream_t st, int ws_slot, cudaStream_t tail, int phase);

static void mark_stage(cb2_ctx* c, int k, cudaStream_t st) {
    if (c->stage_timing && c->stage_ev[k]) cudaEventRecord(c->stage_ev[k], st);
}

// `tail`: the stream the later EPA tiers and the second chance run on (the host pipeline gives them
// their own, so that the next chunk on `st` does not queue behind a long tier-2 tail); by default
// everything is issued on `st`.
static int launch_op(cb2_ctx* c, op_kind op, uint32_t strategy, narrow_args& A, cudaStream_t st,
                     int ws_slot = 0, cudaStream_t tail = nullptr, int phase = 0) {
    if (!tail) tail = st;
    int rc = launch_op_inner(c, op, strategy, A, st, ws_slot, tail, phase);
    if (rc || !((op == OP_INTERSECTION && strategy == CB2_FULL_RESOLUTION) || op == OP_MANIFOLD)) return rc;
    if (phase == 1) return CB2_OK;
    if (std::getenv("CB2_NO_RETRY")) return CB2_OK;  // diagnosis: leave CB2_SCRATCH_OVERFLOW visible
    // redo the rare pairs whose polytope outgrew the fast kernels' scratch
    rc = ensure_retry(c, ws_slot);
    if (rc) return rc;
    cudaError_t e = launch_overflow_retry(A, c->retry[ws_slot], tail);
    c->launches += 2;
    mark_stage(c, 4, tail);
    if (c->stage_timing) {
        c->stage_stream = st;
        c->stage_valid = !c->legacy_epa;
    }
    if (e != cudaSuccess) {
        c->err = std::string("kernel launch: ") + cudaGetErrorString(e);
        return CB2_ERR_CUDA;
    }
    return CB2_OK;
}

static int launch_op_inner(cb2_ctx* c, op_kind op, uint32_t strategy, narrow_args& A,
                           cudaStream_t st, int ws_slot, cudaStream_t tail, int phase) {
    A.T.shapes = c->d_shapes;
    A.T.verts = c->d_verts;
    A.T.cells = c->d_cells;
    A.T.ext = c->d_cell_ext;
    A.n_shapes = c->n_shapes;
    if (!c->d_bad_index) {
        CK(c, cudaMalloc(&c->d_bad_index, sizeof(uint32_t)));
        CK(c, cudaMemset(c->d_bad_index, 0, sizeof(uint32_t)));
        CK(c, cudaMallocHost(&c->h_bad_index, sizeof(uint32_t)));
    }
    A.bad_index = c->d_bad_index;
    cudaError_t e;
    if (op == OP_INTERSECT) {
        e = launch_gjk_intersect(A, st);
        ++c->launches;
        if (e != cudaSuccess) {
            c->err = std::string("kernel launch: ") + cudaGetErrorString(e);
            return CB2_ERR_CUDA;
        }
        return CB2_OK;
    }
    if ((op == OP_INTERSECTION && !c->legacy_epa) || op == OP_MANIFOLD) {
        // GJK (thread per pair) -> compacted hit list -> cooperative EPA tiers
        if (A.n > 0x7FFFFFFFull) return fail(c, CB2_ERR_ARG, "more than 2^31-1 pairs in one batch");
        const bool full = strategy == CB2_FULL_RESOLUTION || op == OP_MANIFOLD;
        int rc = ensure_ws(c, ws_slot, A.n);
        if (rc) return rc;
        const narrow_ws& W = c->ws[ws_slot];
        if (!c->d_queue[ws_slot]) CK(c, cudaMalloc(&c->d_queue[ws_slot], 64));
        A.queue = c->d_queue[ws_slot];
        e = cudaSuccess;
        if (phase != 2) {
            CK(c, cudaMemsetAsync(W.EC, 0, sizeof(epa_counters), st));
            mark_stage(c, 0, st);
            trace_stage("begin", st);
            if (op == OP_MANIFOLD) {
                A.seed_simplex = W.simplex;  // the second-chance kernel restarts from the caller's simplex
                e = launch_manifold_seed(A, W, st);
            } else {
                e = launch_gjk_hits(A, full, W, c->sm_count, st);
            }
            ++c->launches;
            mark_stage(c, 1, st);
            trace_stage("gjk", st);
        } else if (op == OP_MANIFOLD) {
            A.seed_simplex = W.simplex;
        }
        if (e == cudaSuccess && full && phase != 1) {
            e = launch_epa_tier(0, A, W, c->sm_count, st);
            mark_stage(c, 2, st);
            trace_stage("tier0", st);
            if (tail != st && e == cudaSuccess) {
                if (!c->ev_tail[ws_slot])
                    CK(c, cudaEventCreateWithFlags(&c->ev_tail[ws_slot], cudaEventDisableTiming));
                CK(c, cudaEventRecord(c->ev_tail[ws_slot], st));
                CK(c, cudaStreamWaitEvent(tail, c->ev_tail[ws_slot], 0));
            }
            for (int k = 1; k < EPA_TIERS && e == cudaSuccess; ++k) {
                e = launch_epa_tier(k, A, W, c->sm_count, tail);
                trace_stage(k == 1 ? "tier1" : "tier2", tail);
            }
            mark_stage(c, 3, tail);
            c->launches += EPA_TIERS;
        }
        if (e != cudaSuccess) {
            c->err = std::string("kernel launch: ") + cudaGetErrorString(e);
            return CB2_ERR_CUDA;
        }
        return CB2_OK;
    }
    if (op != OP_INTERSECTION) {
        if (!c->d_queue[ws_slot]) CK(c, cudaMalloc(&c->d_queue[ws_slot], 64));
        A.queue = c->d_queue[ws_slot];
    }
    if (op == OP_INTERSECTION) e = launch_intersection(A, strategy == CB2_FULL_RESOLUTION, st);
    else if (op == OP_DISTANCE) e = launch_distance(A, c->sm_count, st);
    else e = launch_time_of_impact(A, c->sm_count, st);
    ++c->launches;
    if (e != cudaSuccess) {
        c->err = std::string("kernel launch: ") + cudaGetErrorString(e);
        return CB2_ERR_CUDA;
    }
    return CB2_OK;
}

static int narrow_dev(cb2_ctx* c, op_kind op, uint32_t strategy, const cb2_settings* settings,
                      const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform* l_t0,
                      const cb2_xform* l_t1, const cb2_xform* r_t0, const cb2_xform* r_t1,
                      const uint32_t* pair_body, uint64_t n, uint32_t iter_cap, cb2_contact* out_c,
                      float* out_d, uint8_t* status, void* stream, cb2_support_point* out_sx = nullptr,
                      const cb2_support_point* in_sx = nullptr, const uint32_t* in_len = nullptr) {
    if (!c) return CB2_ERR_ARG;
    int rc = check_settings(c, settings);
    if (rc) return rc;
    if (op == OP_INTERSECTION && strategy > CB2_COLLISION_ONLY) return fail(c, CB2_ERR_ARG, "bad strategy");
    if (n == 0) return CB2_OK;
    if (!c->d_shapes) return fail(c, CB2_ERR_ARG, "cb2_shapes_upload has not been called");
    if (!l_shape || !l_t0 || !status || (!pair_body && (!r_shape || !r_t0)))
        return fail(c, CB2_ERR_ARG, "null argument");
    if (op == OP_INTERSECT ? !out_sx : (op == OP_DISTANCE ? !out_d : !out_c))
        return fail(c, CB2_ERR_ARG, "null output");
    if (op == OP_MANIFOLD && !in_sx) return fail(c, CB2_ERR_ARG, "null simplex");
    if (op == OP_TOI && (!l_t1 || !r_t1)) return fail(c, CB2_ERR_ARG, "null end transform");
    CK(c, cudaSetDevice(c->device));
    narrow_args A;
    std::memset(&A, 0, sizeof(A));
    A.out_simplex = out_sx;
    A.in_simplex = in_sx;
    A.in_simplex_len = in_len;
    A.l_shape = l_shape;
    A.r_shape = r_shape;
    A.l_t = reinterpret_cast<const float4*>(l_t0);
    A.r_t = reinterpret_cast<const float4*>(pair_body ? l_t0 : r_t0);
    A.l_t1 = reinterpret_cast<const float4*>(l_t1);
    A.r_t1 = reinterpret_cast<const float4*>(r_t1);
    A.pair_body = pair_body;
    A.n = n;
    A.settings = *settings;
    A.iter_cap = iter_cap ? iter_cap : 1024u;
    A.out_contact = out_c;
    A.out_distance = out_d;
    A.out_status = status;
    return launch_op(c, op, strategy, A, static_cast<cudaStream_t>(stream));
}

// ---- host-buffer pipeline ------------------------------------------------------------------------------
// A batch streams through the device in chunks over ROLES, not peers: one copy-in stream that runs
// ahead as far as the four staging buffers allow, three compute streams that take the chunks in
// turn (the low-occupancy tail of chunk k's EPA tiers overlaps chunks k+1, k+2, while chunks still
// finish one after the other), and one copy-out stream.  Staggered completion is the point: with
// four equal peers (copy-in, kernels and copy-out of a chunk all on that chunk's stream) the four
// chunks of a wave finished together and the wave's copies sat outside the compute — CB2_TRACE=1
// showed the SMs idle for ~3 ms of a 24.4 ms call; this arrangement takes the same call to 21.3 ms.
//   in(k):   wait free[buf]  -> H2D          -> record in[buf]
//   comp(k): wait in[buf]    -> kernels      -> record k[buf]
//   out(k):  wait k[buf]     -> D2H          -> record free[buf]
struct host_pipeline {
    cb2_ctx* c;
    cudaStream_t s_in, s_out, s_comp[2];
    static constexpr int NB = cb2_ctx::NS;  // staging buffers
};
static int pipeline_prepare(cb2_ctx* c, size_t stage_bytes) {
    if (c->stage_cap < stage_bytes) {
        for (int i = 0; i < cb2_ctx::NS; ++i) {
            if (c->d_stage[i]) cudaFree(c->d_stage[i]);
            c->d_stage[i] = nullptr;
        }
        c->stage_cap = 0;
        for (int i = 0; i < cb2_ctx::NS; ++i) CK(c, cudaMalloc(&c->d_stage[i], stage_bytes));
        c->stage_cap = stage_bytes;
    }
    if (!c->s_in) CK(c, cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    if (!c->s_out) CK(c, cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < cb2_ctx::NS; ++i) {
        if (!c->ev_in[i]) CK(c, cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
        if (!c->ev_k[i]) CK(c, cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming));
        if (!c->ev_free[i]) CK(c, cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
        if (!c->ev_front[i]) CK(c, cudaEventCreateWithFlags(&c->ev_front[i], cudaEventDisableTiming));
        if (!c->ev_wsfree[i]) CK(c, cudaEventCreateWithFlags(&c->ev_wsfree[i], cudaEventDisableTiming));
    }
    if (!c->s_front) {
        int least = 0, greatest = 0;
        CK(c, cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CK(c, cudaStreamCreateWithPriority(&c->s_front, cudaStreamNonBlocking, greatest));
        // the later tiers sit between the GJK stages and tier 0: they are short, and finishing them
        // releases a staging buffer and a workspace slot
        const int mid = greatest < least ? greatest + 1 : greatest;
        for (int i = 0; i < 2; ++i) CK(c, cudaStreamCreateWithPriority(&c->s_tail[i], cudaStreamNonBlocking, mid));
    }
    return CB2_OK;
}
// chunk sizes: short chunks at both ends (the first copy-in and the last copy-out are the only
// transfers nothing can hide), full chunks in between
static std::vector<uint64_t> pipeline_chunks(uint64_t n, uint64_t chunk, int levels, int levels_tail) {
    std::vector<uint64_t> v;
    if (levels <= 0 && levels_tail <= 0) {
        for (uint64_t b = 0; b < n; b += chunk) v.push_back(n - b < chunk ? n - b : chunk);
        return v;
    }
    if (n <= chunk) {
        v.push_back(n);
        return v;
    }
    std::vector<uint64_t> ramp, ramp_t;  // chunk / 2^levels ... chunk / 2
    for (int k = levels; k >= 1; --k) ramp.push_back(chunk >> k);
    for (int k = levels_tail; k >= 1; --k) ramp_t.push_back(chunk >> k);
    uint64_t head = 0, tail = 0;
    if (n >= 4 * chunk) {
        for (uint64_t r : ramp) {
            v.push_back(r);
            head += r;
        }
        for (uint64_t r : ramp_t) tail += r;
    }
    // the middle in equal chunks (no short remainder chunk: every chunk pays the latency of its tiers' tails)
    uint64_t left = n - head - tail;
    uint64_t parts = (left + chunk - 1) / chunk;
    while (left) {
        const uint64_t m = (left + parts - 1) / parts;
        v.push_back(m);
        left -= m;
        --parts;
    }
    if (tail)
        for (size_t k = ramp_t.size(); k-- > 0;) v.push_back(ramp_t[k]);
    return v;
}
template <class FIn, class FComp, class FOut>
static int pipeline_run(cb2_ctx* c, uint64_t n, FIn copy_in, FComp compute, FOut copy_out,
                        bool split_tail = false) {
    static const bool trace = std::getenv("CB2_TRACE") && std::getenv("CB2_TRACE")[0] == '1';
    cudaStream_t s_in = c->s_in, s_out = c->s_out;
    const std::vector<uint64_t> sizes = pipeline_chunks(n, c->chunk_pairs, c->ramp, c->ramp_tail);
    std::vector<cudaEvent_t> tev;
    cudaEvent_t t0 = nullptr;
    if (trace) {
        cudaEventCreate(&t0);
        cudaEventRecord(t0, s_in);
    }
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        tev.push_back(e);
    };
    uint64_t b = 0;
    int rc = CB2_OK;
    // inside the loop a failed CUDA call must not return: copies into the caller's buffers may
    // still be in flight, and every exit has to pass the stream synchronisation below
#define CKB(x)                                                             \
    {                                                                      \
        cudaError_t e_ = (x);                                              \
        if (e_ != cudaSuccess) {                                           \
            c->err = std::string(#x) + ": " + cudaGetErrorString(e_);      \
            rc = CB2_ERR_CUDA;                                             \
            break;                                                         \
        }                                                                  \
    }
    for (size_t k = 0; k < sizes.size() && rc == CB2_OK; ++k) {
        const uint64_t m = sizes[k];
        // split_tail (the GJK -> EPA pipeline): two front streams carry GJK + tier 0 of alternating
        // chunks, two tail streams the later tiers and the second chance; the workspace slot follows
        // the staging buffer, whose `free` event already orders its reuse
        const int buf = (int)(k % cb2_ctx::NS);
        const int slot = split_tail ? buf : (int)(k % (size_t)c->n_cstreams);
        char* base = c->d_stage[buf];
        cudaStream_t s_comp = split_tail ? c->stream[k % 2] : c->stream[slot];
        cudaStream_t s_tail = split_tail ? c->stream[2 + k % 2] : s_comp;
        if (k >= (size_t)cb2_ctx::NS) CKB(cudaStreamWaitEvent(s_in, c->ev_free[buf], 0))
        rc = copy_in(b, m, base, s_in);
        if (rc) break;
        CKB(cudaEventRecord(c->ev_in[buf], s_in))
        mark(s_in);
        CKB(cudaStreamWaitEvent(s_comp, c->ev_in[buf], 0))
        rc = compute(b, m, base, s_comp, slot, s_tail);
        if (rc) break;
        CKB(cudaEventRecord(c->ev_k[buf], s_tail))
        mark(s_tail);
        CKB(cudaStreamWaitEvent(s_out, c->ev_k[buf], 0))
        rc = copy_out(b, m, base, s_out);
        if (rc) break;
        CKB(cudaEventRecord(c->ev_free[buf], s_out))
        mark(s_out);
        b += m;
    }
#undef CKB
    for (int i = 0; i < cb2_ctx::NS + 2; ++i) {
        cudaError_t e = cudaStreamSynchronize(i < cb2_ctx::NS ? c->stream[i] : (i == cb2_ctx::NS ? s_in : s_out));
        if (e != cudaSuccess && rc == CB2_OK) {
            c->err = std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e);
            rc = CB2_ERR_CUDA;
        }
    }
    if (trace) {
        for (size_t q = 0; q + 2 < tev.size(); q += 3) {
            float a_ = 0, b_ = 0, d_ = 0;
            cudaEventElapsedTime(&a_, t0, tev[q]);
            cudaEventElapsedTime(&b_, t0, tev[q + 1]);
            cudaEventElapsedTime(&d_, t0, tev[q + 2]);
            std::fprintf(stderr, "[cb2 trace] chunk %zu (%llu pairs): in %.3f  kernels %.3f  out %.3f ms\n", q / 3,
                         (unsigned long long)sizes[q / 3], a_, b_, d_);
        }
        std::fprintf(stderr, "[cb2 trace] stages:");
        for (auto& m : g_trace) {
            float t_ = 0;
            cudaEventElapsedTime(&t_, t0, m.ev);
            std::fprintf(stderr, "%s %s %.3f", std::strcmp(m.what, "begin") == 0 ? "\n[cb2 trace]  " : ",", m.what, t_);
            cudaEventDestroy(m.ev);
        }
        std::fprintf(stderr, "\n");
        g_trace.clear();
        for (auto e : tev) cudaEventDestroy(e);
        cudaEventDestroy(t0);
    }
    return rc;
}
// The GJK-ahead variant (FullResolution intersections): a chunk's GJK stage cannot share the SMs
// with the previous chunk's persistent EPA tier 0 (registers), so on one stream per chunk it ran in
// that kernel's tail and the chunk's own tier 0 waited for it — 2.8 ms per 512 K pairs against 2.34 ms
// inside one resident batch.  Here the GJK stages of all chunks go to ONE high-priority stream and
// run as soon as their inputs have landed, i.e. one chunk ahead of the tier 0 that is about to
// start: when tier 0 of chunk k-1 drains, GJK(k+1) and tier 0 of chunk k (whose own GJK finished a
// chunk earlier) are both ready, the block scheduler serves the high-priority GJK blocks first and
// tier 0's persistent blocks take what is left, then the rest.  No kernel ever waits for another
// that still has to be scheduled behind it.  NW workspace slots (hit list, simplex hand-off,
// dumps) rotate under their own `free` events; the staging buffers (NS) live from the copy-in to
// the copy-out of a chunk, about five chunk periods.
template <class FIn, class FComp, class FOut>
static int pipeline_run_ahead(cb2_ctx* c, uint64_t n, int nw, FIn copy_in, FComp compute, FOut copy_out) {
    static const bool trace = std::getenv("CB2_TRACE") && std::getenv("CB2_TRACE")[0] == '1';
    cudaStream_t s_in = c->s_in, s_out = c->s_out, s_front = c->s_front;
    const std::vector<uint64_t> sizes = pipeline_chunks(n, c->chunk_pairs, c->ramp, c->ramp_tail);
    std::vector<uint64_t> starts(sizes.size());
    {
        uint64_t b = 0;
        for (size_t k = 0; k < sizes.size(); ++k) {
            starts[k] = b;
            b += sizes[k];
        }
    }
    cudaEvent_t t0 = nullptr;
    std::vector<cudaEvent_t> tev;
    if (trace) {
        cudaEventCreate(&t0);
        cudaEventRecord(t0, s_in);
    }
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        tev.push_back(e);
    };
    int rc = CB2_OK;
#define CKA(x)                                                             \
    {                                                                      \
        cudaError_t e_ = (x);                                              \
        if (e_ != cudaSuccess && rc == CB2_OK) {                           \
            c->err = std::string(#x) + ": " + cudaGetErrorString(e_);      \
            rc = CB2_ERR_CUDA;                                             \
        }                                                                  \
    }
    // EPA tiers + second chance + copy-out of chunk j
    auto issue_back = [&](size_t j) {
        const int buf = (int)(j % cb2_ctx::NS), wslot = (int)(j % (size_t)nw);
        cudaStream_t sb = c->stream[j % (size_t)c->n_cstreams];
        char* base = c->d_stage[buf];
        cudaStream_t stl = c->tail_prio ? c->s_tail[j % 2] : sb;
        CKA(cudaStreamWaitEvent(sb, c->ev_front[wslot], 0))
        if (rc == CB2_OK) rc = compute(starts[j], sizes[j], base, sb, wslot, stl, 2);
        CKA(cudaEventRecord(c->ev_k[buf], stl))
        CKA(cudaEventRecord(c->ev_wsfree[wslot], stl))
        mark(stl);
        CKA(cudaStreamWaitEvent(s_out, c->ev_k[buf], 0))
        if (rc == CB2_OK) rc = copy_out(starts[j], sizes[j], base, s_out);
        CKA(cudaEventRecord(c->ev_free[buf], s_out))
        mark(s_out);
    };
    size_t issued_back = 0;
    for (size_t k = 0; k < sizes.size() && rc == CB2_OK; ++k) {
        const int buf = (int)(k % cb2_ctx::NS), wslot = (int)(k % (size_t)nw);
        char* base = c->d_stage[buf];
        // the staging buffer and the workspace slot of chunk k were last used by chunks k-NS and
        // k-nw: their back parts must have been ISSUED before anybody can wait for their events
        while (issued_back + (size_t)cb2_ctx::NS <= k || issued_back + (size_t)nw <= k) issue_back(issued_back++);
        if (rc != CB2_OK) break;
        if (k >= (size_t)cb2_ctx::NS) CKA(cudaStreamWaitEvent(s_in, c->ev_free[buf], 0))
        if (rc == CB2_OK) rc = copy_in(starts[k], sizes[k], base, s_in);
        CKA(cudaEventRecord(c->ev_in[buf], s_in))
        mark(s_in);
        CKA(cudaStreamWaitEvent(s_front, c->ev_in[buf], 0))
        if (k >= (size_t)nw) CKA(cudaStreamWaitEvent(s_front, c->ev_wsfree[wslot], 0))
        if (rc == CB2_OK) rc = compute(starts[k], sizes[k], base, s_front, wslot, s_front, 1);
        CKA(cudaEventRecord(c->ev_front[wslot], s_front))
        // GJK(k) precedes the EPA tiers of chunk k-1 in issue order as well
        while (issued_back + 1 <= k && rc == CB2_OK) issue_back(issued_back++);
    }
    // every chunk whose GJK stage was issued gets its back part (a failed call still drains cleanly)
    while (issued_back < sizes.size() && rc == CB2_OK) issue_back(issued_back++);
#undef CKA
    for (int i = 0; i < cb2_ctx::NS + 5; ++i) {
        cudaStream_t st = i < cb2_ctx::NS ? c->stream[i]
                          : (i == cb2_ctx::NS ? s_in : (i == cb2_ctx::NS + 1 ? s_out : (i == cb2_ctx::NS + 2 ? s_front : c->s_tail[i - cb2_ctx::NS - 3])));
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess && rc == CB2_OK) {
            c->err = std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e);
            rc = CB2_ERR_CUDA;
        }
    }
    if (trace) {
        // per chunk: copy-in | (kernels, copy-out) are recorded in issue order: in(k) first, then the
        // back of chunk k-1 — print them as they come
        for (size_t q = 0; q < tev.size(); ++q) {
            float a_ = 0;
            cudaEventElapsedTime(&a_, t0, tev[q]);
            std::fprintf(stderr, "[cb2 trace] mark %zu: %.3f ms\n", q, a_);
        }
        std::fprintf(stderr, "[cb2 trace] stages:");
        for (auto& m : g_trace) {
            float t_ = 0;
            cudaEventElapsedTime(&t_, t0, m.ev);
            std::fprintf(stderr, "%s %s %.3f", std::strcmp(m.what, "begin") == 0 ? "\n[cb2 trace]  " : ",", m.what, t_);
            cudaEventDestroy(m.ev);
        }
        std::fprintf(stderr, "\n");
        g_trace.clear();
        for (auto e : tev) cudaEventDestroy(e);
        cudaEventDestroy(t0);
    }
    return rc;
}
// the 2-D host path checks its shape indices on the host, chunk by chunk (the 3-D kernels check on
// the device, see narrow_host)
static bool indices_in_range(const uint32_t* a, const uint32_t* b, uint64_t m, uint32_t limit) {
    uint32_t worst = 0;
    for (uint64_t i = 0; i < m; ++i) {
        worst = a[i] > worst ? a[i] : worst;
        worst = b[i] > worst ? b[i] : worst;
    }
    return worst < limit;
}

// Host-buffer path of the 3-D narrow phase.
// `pair_body` (intersection only): the pairs index ONE body table — l_shape / l_t0 then hold n_bodies
// shape indices / transforms, uploaded once per call, and only the 8 bytes of a pair's two body
// indices are staged per pair (72 bytes per pair otherwise).
static int narrow_host(cb2_ctx* c, op_kind op, uint32_t strategy, const cb2_settings* settings,
                       const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform* l_t0,
                       const cb2_xform* l_t1, const cb2_xform* r_t0, const cb2_xform* r_t1,
                       uint64_t n, uint32_t iter_cap, cb2_contact* out_c, float* out_d,
                       uint8_t* status, const uint32_t* pair_body = nullptr, uint64_t n_bodies = 0) {
    if (!c) return CB2_ERR_ARG;
    int rc = check_settings(c, settings);
    if (rc) return rc;
    if (op == OP_INTERSECTION && strategy > CB2_COLLISION_ONLY) return fail(c, CB2_ERR_ARG, "bad strategy");
    if (n == 0) return CB2_OK;
    if (!c->d_shapes) return fail(c, CB2_ERR_ARG, "cb2_shapes_upload has not been called");
    const bool bodies = pair_body != nullptr;
    if (!l_shape || !l_t0 || !status || (!bodies && (!r_shape || !r_t0))) return fail(c, CB2_ERR_ARG, "null argument");
    if (bodies && (op != OP_INTERSECTION || n_bodies == 0)) return fail(c, CB2_ERR_ARG, "empty body table");
    if (op == OP_DISTANCE ? !out_d : !out_c) return fail(c, CB2_ERR_ARG, "null output");
    if (op == OP_TOI && (!l_t1 || !r_t1)) return fail(c, CB2_ERR_ARG, "null end transform");
    CK(c, cudaSetDevice(c->device));
    const uint64_t chunk = n < c->chunk_pairs ? n : c->chunk_pairs;
    const int n_xf = bodies ? 0 : (op == OP_TOI ? 4 : 2);
    const size_t out_rec = op == OP_DISTANCE ? sizeof(float) : sizeof(cb2_contact);
    // staging layout per buffer: xforms | l_shape | r_shape | out | status   (256-byte aligned)
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t xf_bytes = al(chunk * sizeof(cb2_xform));
    const size_t o_ls = n_xf * xf_bytes;
    const size_t o_rs = o_ls + al(chunk * 4);  // (bodies: the 8-byte index pairs start at o_ls and run over o_rs)
    const size_t o_out = o_rs + al(chunk * 4) + (bodies ? 256 : 0);
    const size_t o_st = o_out + al(chunk * out_rec);
    rc = pipeline_prepare(c, o_st + al(chunk));
    if (rc) return rc;
    // size every compute slot's workspace for the largest chunk now: growing one in mid-pipeline
    // (the ramp starts with short chunks) would free buffers under a device-wide synchronisation
    // CB2_SPLIT_TAIL=1 (experiment, measured no gain: 22.6 against 22.1 ms per 4 M-pair call): front /
    // tail streams and one workspace per staging buffer, so that a chunk's GJK does not queue behind
    // the tier-2 tail of the chunk three before it.  Default: every stage of a chunk on one of
    // n_cstreams streams.
    static const bool split_env = std::getenv("CB2_SPLIT_TAIL") && std::getenv("CB2_SPLIT_TAIL")[0] == '1';
    const bool split = split_env && op == OP_INTERSECTION && !c->legacy_epa && strategy == CB2_FULL_RESOLUTION &&
                       cb2_ctx::NS >= 4;
    // CB2_AHEAD (default on): the GJK stages run ahead on their own high-priority stream
    // (pipeline_run_ahead); nw workspace slots rotate
    const bool ahead = c->ahead != 0 && !split && op == OP_INTERSECTION && !c->legacy_epa &&
                       strategy == CB2_FULL_RESOLUTION && cb2_ctx::NS >= 6;
    const int nw = ahead ? (c->n_cstreams + 1 < cb2_ctx::NS ? c->n_cstreams + 1 : cb2_ctx::NS) : 0;
    if (op == OP_INTERSECTION && !c->legacy_epa)
        for (int slot = 0; slot < (split ? cb2_ctx::NS : (ahead ? nw : c->n_cstreams)); ++slot) {
            rc = ensure_ws(c, slot, chunk);
            if (rc) return rc;
            if (ahead) {
                rc = ensure_retry(c, slot);
                if (rc) return rc;
            }
        }
    if (bodies) {
        if (c->body_cap < n_bodies) {
            if (c->d_body_shape) cudaFree(c->d_body_shape);
            if (c->d_body_t) cudaFree(c->d_body_t);
            c->d_body_shape = nullptr;
            c->d_body_t = nullptr;
            c->body_cap = 0;
            CK(c, cudaMalloc(&c->d_body_shape, n_bodies * sizeof(uint32_t)));
            CK(c, cudaMalloc(&c->d_body_t, n_bodies * sizeof(cb2_xform)));
            c->body_cap = n_bodies;
        }
        // on the copy-in stream: the first chunk's `inputs landed` event is recorded behind them
        CK(c, cudaMemcpyAsync(c->d_body_shape, l_shape, n_bodies * sizeof(uint32_t), cudaMemcpyHostToDevice, c->s_in));
        CK(c, cudaMemcpyAsync(c->d_body_t, l_t0, n_bodies * sizeof(cb2_xform), cudaMemcpyHostToDevice, c->s_in));
    }
    const cb2_xform* xf_src[4] = {l_t0, r_t0, l_t1, r_t1};
    // a Rust slice index would panic on an out-of-range shape index: the kernels check every index
    // they load (status CB2_BAD_INDEX, nothing is read through it) and raise a device flag; the call
    // then fails as a whole (outputs of a failed call are unspecified)
    if (c->d_bad_index) CK(c, cudaMemsetAsync(c->d_bad_index, 0, sizeof(uint32_t), c->s_in));
    auto copy_in = [&](uint64_t b, uint64_t m, char* base, cudaStream_t st) -> int {
            for (int x = 0; x < n_xf; ++x)
                CK(c, cudaMemcpyAsync(base + x * xf_bytes, xf_src[x] + b, m * sizeof(cb2_xform),
                                      cudaMemcpyHostToDevice, st));
            if (bodies) {
                CK(c, cudaMemcpyAsync(base + o_ls, pair_body + 2 * b, m * 8, cudaMemcpyHostToDevice, st));
                return CB2_OK;
            }
            CK(c, cudaMemcpyAsync(base + o_ls, l_shape + b, m * 4, cudaMemcpyHostToDevice, st));
            CK(c, cudaMemcpyAsync(base + o_rs, r_shape + b, m * 4, cudaMemcpyHostToDevice, st));
            return CB2_OK;
        };
    auto compute_phase = [&](uint64_t, uint64_t m, char* base, cudaStream_t st, int slot, cudaStream_t tail,
                             int phase) -> int {
            narrow_args A;
            std::memset(&A, 0, sizeof(A));
            A.l_shape = reinterpret_cast<const uint32_t*>(base + o_ls);
            A.r_shape = reinterpret_cast<const uint32_t*>(base + o_rs);
            A.l_t = reinterpret_cast<const float4*>(base);
            A.r_t = reinterpret_cast<const float4*>(base + xf_bytes);
            if (bodies) {
                A.pair_body = reinterpret_cast<const uint32_t*>(base + o_ls);
                A.l_shape = c->d_body_shape;
                A.r_shape = nullptr;
                A.l_t = A.r_t = c->d_body_t;
                A.n_bodies = n_bodies;
            }
            if (op == OP_TOI) {
                A.l_t1 = reinterpret_cast<const float4*>(base + 2 * xf_bytes);
                A.r_t1 = reinterpret_cast<const float4*>(base + 3 * xf_bytes);
            }
            A.n = m;
            A.settings = *settings;
            A.iter_cap = iter_cap ? iter_cap : 1024u;
            A.out_contact = reinterpret_cast<cb2_contact*>(base + o_out);
            A.out_distance = reinterpret_cast<float*>(base + o_out);
            A.out_status = reinterpret_cast<uint8_t*>(base + o_st);
            return launch_op(c, op, strategy, A, st, slot, tail, phase);
        };
    auto copy_out = [&](uint64_t b, uint64_t m, char* base, cudaStream_t st) -> int {
            if (op == OP_DISTANCE)
                CK(c, cudaMemcpyAsync(out_d + b, base + o_out, m * sizeof(float), cudaMemcpyDeviceToHost, st));
            else
                CK(c, cudaMemcpyAsync(out_c + b, base + o_out, m * sizeof(cb2_contact),
                                      cudaMemcpyDeviceToHost, st));
            CK(c, cudaMemcpyAsync(status + b, base + o_st, m, cudaMemcpyDeviceToHost, st));
            return CB2_OK;
        };
    if (ahead)
        rc = pipeline_run_ahead(
            c, n, nw, copy_in,
            [&](uint64_t b, uint64_t m, char* base, cudaStream_t st, int slot, cudaStream_t tail, int phase) -> int {
                return compute_phase(b, m, base, st, slot, tail, phase);
            },
            copy_out);
    else
        rc = pipeline_run(
            c, n, copy_in,
            [&](uint64_t b, uint64_t m, char* base, cudaStream_t st, int slot, cudaStream_t tail) -> int {
                return compute_phase(b, m, base, st, slot, tail, 0);
            },
            copy_out, split);
    if (rc == CB2_OK && c->d_bad_index) {
        CK(c, cudaMemcpy(c->h_bad_index, c->d_bad_index, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (*c->h_bad_index) return fail(c, CB2_ERR_ARG, bodies ? "shape or body index out of range" : "shape index out of range");
    }
    return rc;
}

static int ensure_misc(cb2_ctx* c, size_t bytes) {
    if (c->misc_cap >= bytes) return CB2_OK;
    if (c->d_misc) cudaFree(c->d_misc);
    c->d_misc = nullptr;
    c->misc_cap = 0;
    CK(c, cudaMalloc(&c->d_misc, bytes));
    c->misc_cap = bytes;
    return CB2_OK;
}

extern "C" {

int cb2_gjk3_intersection(cb2_ctx* c, uint32_t strategy, const cb2_settings* settings,
                          const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform* l_t,
                          const cb2_xform* r_t, uint64_t n, cb2_contact* out, uint8_t* status) {
    return narrow_host(c, OP_INTERSECTION, strategy, settings, l_shape, r_shape, l_t, nullptr, r_t,
                       nullptr, n, 0, out, nullptr, status);
}
int cb2_gjk3_distance(cb2_ctx* c, const cb2_settings* settings, const uint32_t* l_shape,
                      const uint32_t* r_shape, const cb2_xform* l_t, const cb2_xform* r_t,
                      uint64_t n, float* out, uint8_t* status) {
    return narrow_host(c, OP_DISTANCE, 0, settings, l_shape, r_shape, l_t, nullptr, r_t, nullptr, n,
                       0, nullptr, out, status);
}
int cb2_gjk3_time_of_impact(cb2_ctx* c, const cb2_settings* settings, const uint32_t* l_shape,
                            const uint32_t* r_shape, const cb2_xform* l_t0, const cb2_xform* l_t1,
                            const cb2_xform* r_t0, const cb2_xform* r_t1, uint64_t n,
                            uint32_t iter_cap, cb2_contact* out, uint8_t* status) {
    return narrow_host(c, OP_TOI, 0, settings, l_shape, r_shape, l_t0, l_t1, r_t0, r_t1, n, iter_cap,
                       out, nullptr, status);
}

int cb2_gjk3_intersect_dev(cb2_ctx* c, const cb2_settings* settings, const uint32_t* l_shape,
                           const uint32_t* r_shape, const cb2_xform* l_t, const cb2_xform* r_t,
                           uint64_t n, cb2_support_point* simplex_out, uint8_t* status, void* stream) {
    return narrow_dev(c, OP_INTERSECT, 0, settings, l_shape, r_shape, l_t, nullptr, r_t, nullptr, nullptr,
                      n, 0, nullptr, nullptr, status, stream, simplex_out);
}
int cb2_gjk3_get_contact_manifold_dev(cb2_ctx* c, const cb2_settings* settings,
                                      const cb2_support_point* simplex, const uint32_t* simplex_len,
                                      const uint32_t* l_shape, const uint32_t* r_shape,
                                      const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                                      cb2_contact* out, uint8_t* status, void* stream) {
    return narrow_dev(c, OP_MANIFOLD, 0, settings, l_shape, r_shape, l_t, nullptr, r_t, nullptr, nullptr,
                      n, 0, out, nullptr, status, stream, nullptr, simplex, simplex_len);
}

// Host-buffer versions of the two: one staged batch (these are the building blocks a caller uses
// when it wants the simplex between the two halves of GJK::intersection; the chunked pipeline of
// cb2_gjk3_intersection is the fast path for the whole thing).
static int simplex_host(cb2_ctx* c, bool manifold, const cb2_settings* settings,
                        const cb2_support_point* in_sx, const uint32_t* in_len, const uint32_t* l_shape,
                        const uint32_t* r_shape, const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                        cb2_support_point* out_sx, cb2_contact* out_c, uint8_t* status) {
    if (!c) return CB2_ERR_ARG;
    if (n == 0) return CB2_OK;
    if (!l_shape || !r_shape || !l_t || !r_t || !status || (manifold ? (!in_sx || !out_c) : !out_sx))
        return fail(c, CB2_ERR_ARG, "null argument");
    CK(c, cudaSetDevice(c->device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_lt = 0, o_rt = al(n * sizeof(cb2_xform)), o_ls = o_rt + al(n * sizeof(cb2_xform)),
                 o_rs = o_ls + al(n * 4), o_sx = o_rs + al(n * 4),
                 o_len = o_sx + al(n * 4 * sizeof(cb2_support_point)), o_out = o_len + al(n * 4),
                 o_st = o_out + al(n * sizeof(cb2_contact));
    int rc = ensure_misc(c, o_st + al(n));
    if (rc) return rc;
    char* base = static_cast<char*>(c->d_misc);
    cudaStream_t st = c->stream[0];
    CK(c, cudaMemcpyAsync(base + o_lt, l_t, n * sizeof(cb2_xform), cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(base + o_rt, r_t, n * sizeof(cb2_xform), cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(base + o_ls, l_shape, n * 4, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(base + o_rs, r_shape, n * 4, cudaMemcpyHostToDevice, st));
    if (manifold) {
        CK(c, cudaMemcpyAsync(base + o_sx, in_sx, n * 4 * sizeof(cb2_support_point), cudaMemcpyHostToDevice, st));
        if (in_len) CK(c, cudaMemcpyAsync(base + o_len, in_len, n * 4, cudaMemcpyHostToDevice, st));
    }
    if (c->d_bad_index) CK(c, cudaMemsetAsync(c->d_bad_index, 0, sizeof(uint32_t), st));
    auto* d_sx = reinterpret_cast<cb2_support_point*>(base + o_sx);
    auto* d_ls = reinterpret_cast<uint32_t*>(base + o_ls);
    auto* d_rs = reinterpret_cast<uint32_t*>(base + o_rs);
    auto* d_lt = reinterpret_cast<cb2_xform*>(base + o_lt);
    auto* d_rt = reinterpret_cast<cb2_xform*>(base + o_rt);
    auto* d_st = reinterpret_cast<uint8_t*>(base + o_st);
    if (manifold)
        rc = cb2_gjk3_get_contact_manifold_dev(c, settings, d_sx,
                                               in_len ? reinterpret_cast<uint32_t*>(base + o_len) : nullptr,
                                               d_ls, d_rs, d_lt, d_rt, n,
                                               reinterpret_cast<cb2_contact*>(base + o_out), d_st, st);
    else
        rc = cb2_gjk3_intersect_dev(c, settings, d_ls, d_rs, d_lt, d_rt, n, d_sx, d_st, st);
    if (rc) return rc;
    if (manifold)
        CK(c, cudaMemcpyAsync(out_c, base + o_out, n * sizeof(cb2_contact), cudaMemcpyDeviceToHost, st));
    else
        CK(c, cudaMemcpyAsync(out_sx, base + o_sx, n * 4 * sizeof(cb2_support_point), cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(status, base + o_st, n, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(c->h_bad_index, c->d_bad_index, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    if (*c->h_bad_index) return fail(c, CB2_ERR_ARG, "shape index out of range");
    return CB2_OK;
}
int cb2_gjk3_intersect(cb2_ctx* c, const cb2_settings* settings, const uint32_t* l_shape,
                       const uint32_t* r_shape, const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                       cb2_support_point* simplex_out, uint8_t* status) {
    return simplex_host(c, false, settings, nullptr, nullptr, l_shape, r_shape, l_t, r_t, n, simplex_out,
                        nullptr, status);
}
int cb2_gjk3_get_contact_manifold(cb2_ctx* c, const cb2_settings* settings,
                                  const cb2_support_point* simplex, const uint32_t* simplex_len,
                                  const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform* l_t,
                                  const cb2_xform* r_t, uint64_t n, cb2_contact* out, uint8_t* status) {
    return simplex_host(c, true, settings, simplex, simplex_len, l_shape, r_shape, l_t, r_t, n, nullptr, out,
                        status);
}

int cb2_gjk3_intersection_dev(cb2_ctx* c, uint32_t strategy, const cb2_settings* settings,
                              const uint32_t* l_shape, const uint32_t* r_shape,
                              const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                              cb2_contact* out, uint8_t* status, void* stream) {
    return narrow_dev(c, OP_INTERSECTION, strategy, settings, l_shape, r_shape, l_t, nullptr, r_t,
                      nullptr, nullptr, n, 0, out, nullptr, status, stream);
}
int cb2_gjk3_distance_dev(cb2_ctx* c, const cb2_settings* settings, const uint32_t* l_shape,
                          const uint32_t* r_shape, const cb2_xform* l_t, const cb2_xform* r_t,
                          uint64_t n, float* out, uint8_t* status, void* stream) {
    return narrow_dev(c, OP_DISTANCE, 0, settings, l_shape, r_shape, l_t, nullptr, r_t, nullptr,
                      nullptr, n, 0, nullptr, out, status, stream);
}
int cb2_gjk3_time_of_impact_dev(cb2_ctx* c, const cb2_settings* settings, const uint32_t* l_shape,
                                const uint32_t* r_shape, const cb2_xform* l_t0,
                                const cb2_xform* l_t1, const cb2_xform* r_t0,
                                const cb2_xform* r_t1, uint64_t n, uint32_t iter_cap,
                                cb2_contact* out, uint8_t* status, void* stream) {
    return narrow_dev(c, OP_TOI, 0, settings, l_shape, r_shape, l_t0, l_t1, r_t0, r_t1, nullptr, n,
                      iter_cap, out, nullptr, status, stream);
}
int cb2_gjk3_intersection_bodies_dev(cb2_ctx* c, uint32_t strategy, const cb2_settings* settings,
                                     const uint32_t* body_shape, const cb2_xform* body_t,
                                     const uint32_t* pair_body, uint64_t n, cb2_contact* out,
                                     uint8_t* status, void* stream) {
    if (!pair_body && n) return fail(c, CB2_ERR_ARG, "pair_body is null");
    return narrow_dev(c, OP_INTERSECTION, strategy, settings, body_shape, nullptr, body_t, nullptr,
                      nullptr, nullptr, pair_body, n, 0, out, nullptr, status, stream);
}

int cb2_gjk3_intersection_bodies(cb2_ctx* c, uint32_t strategy, const cb2_settings* settings,
                                 const uint32_t* body_shape, const cb2_xform* body_t, uint64_t n_bodies,
                                 const uint32_t* pair_body, uint64_t n, cb2_contact* out, uint8_t* status) {
    if (!pair_body && n) return fail(c, CB2_ERR_ARG, "pair_body is null");
    return narrow_host(c, OP_INTERSECTION, strategy, settings, body_shape, nullptr, body_t, nullptr, nullptr,
                       nullptr, n, 0, out, nullptr, status, pair_body, n_bodies);
}

// ---- broad phase -------------------------------------------------------------------------------------
int cb2_sap3_find_collider_pairs_dev(cb2_ctx* c, const cb2_aabb3* aabbs, uint64_t n, uint32_t axis,
                                     uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap,
                                     uint64_t* n_pairs, void* stream) {
    if (!c || !n_pairs) return fail(c, CB2_ERR_ARG, "null argument");
    *n_pairs = 0;
    if (axis > 2) return fail(c, CB2_ERR_ARG, "sweep axis must be 0, 1 or 2");
    if (n == 0) return CB2_OK;
    if (!aabbs || !perm_out || (!pairs_out && cap)) return fail(c, CB2_ERR_ARG, "null argument");
    CK(c, cudaSetDevice(c->device));
    const char* err = "";
    int rc = sap_find_pairs(c->sap, aabbs, n, axis, perm_out, pairs_out, cap, n_pairs,
                            static_cast<cudaStream_t>(stream), &c->launches, &err);
    if (rc == CB2_ERR_NOSPC) c->err = "pairs_out too small; *n_pairs holds the required capacity";
    else if (rc) c->err = err;
    return rc;
}

// The pair SET without the reference's Vec order (north_star: "a bit-exact match to the reference's
// pair set"): for callers that feed the pairs straight to the narrow phase, where order is
// irrelevant — the final 40-bit radix sort of the pair keys (a quarter of the call) is skipped.
int cb2_sap3_find_collider_pairs_unordered_dev(cb2_ctx* c, const cb2_aabb3* aabbs, uint64_t n,
                                               uint32_t axis, uint32_t* perm_out, uint32_t* pairs_out,
                                               uint64_t cap, uint64_t* n_pairs, void* stream) {
    if (!c || !n_pairs) return fail(c, CB2_ERR_ARG, "null argument");
    *n_pairs = 0;
    if (axis > 2) return fail(c, CB2_ERR_ARG, "sweep axis must be 0, 1 or 2");
    if (n == 0) return CB2_OK;
    if (!aabbs || !perm_out || (!pairs_out && cap)) return fail(c, CB2_ERR_ARG, "null argument");
    CK(c, cudaSetDevice(c->device));
    const char* err = "";
    int rc = sap_find_pairs(c->sap, aabbs, n, axis, perm_out, pairs_out, cap, n_pairs,
                            static_cast<cudaStream_t>(stream), &c->launches, &err, false, 0u, 0xFFFFFFFFu,
                            true);
    if (rc == CB2_ERR_NOSPC) c->err = "pairs_out too small; *n_pairs holds the required capacity";
    else if (rc) c->err = err;
    return rc;
}

// One GPU's shard of the pair list: every GPU sorts the whole table (the exact stable
// permutation, 24 MB at 1 M boxes) and sweeps it, but reports only the pairs whose later box sits at
// a sorted position in shard `shard` of `n_shards` equal position ranges.  The reference's Vec is
// ordered by that position, so the shards concatenated in rank order ARE the reference's Vec.
int cb2_sap3_find_collider_pairs_shard_dev(cb2_ctx* c, const cb2_aabb3* aabbs, uint64_t n,
                                           uint32_t axis, uint32_t shard, uint32_t n_shards,
                                           uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap,
                                           uint64_t* n_pairs, void* stream) {
    if (!c || !n_pairs) return fail(c, CB2_ERR_ARG, "null argument");
    *n_pairs = 0;
    if (axis > 2) return fail(c, CB2_ERR_ARG, "sweep axis must be 0, 1 or 2");
    if (n_shards == 0 || shard >= n_shards) return fail(c, CB2_ERR_ARG, "shard out of range");
    if (n == 0) return CB2_OK;
    if (!aabbs || !perm_out || (!pairs_out && cap)) return fail(c, CB2_ERR_ARG, "null argument");
    // the same split as collision_rs_b200/sharding.py shard_range(): the first n % n_shards ranks
    // hold one more position
    const uint64_t per = n / n_shards, rem = n % n_shards;
    const uint64_t lo = shard * per + (shard < rem ? shard : rem);
    const uint64_t hi = lo + per + (shard < rem ? 1 : 0);
    CK(c, cudaSetDevice(c->device));
    const char* err = "";
    int rc = sap_find_pairs(c->sap, aabbs, n, axis, perm_out, pairs_out, cap, n_pairs,
                            static_cast<cudaStream_t>(stream), &c->launches, &err, false,
                            (uint32_t)lo, shard + 1 == n_shards ? 0xFFFFFFFFu : (uint32_t)hi);
    if (rc == CB2_ERR_NOSPC) c->err = "pairs_out too small; *n_pairs holds the required capacity";
    else if (rc) c->err = err;
    return rc;
}

// SweepAndPrune2::find_collider_pairs (sweep_prune.rs:18, 76-136 with Variance2 :171-224): the 2-D
// boxes are lifted to z in [0, 1] — the strict z test 1 > 0 && 0 < 1 always holds — and go through
// the 3-D path; Variance2 discards its sums exactly like Variance3, so the next axis is 0 as well.
int cb2_sap2_find_collider_pairs(cb2_ctx* c, const cb2_aabb2* aabbs, uint64_t n, uint32_t* axis_inout,
                                 uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap,
                                 uint64_t* n_pairs) {
    if (!c || !axis_inout || !n_pairs) return fail(c, CB2_ERR_ARG, "null argument");
    if (*axis_inout > 1) return fail(c, CB2_ERR_ARG, "sweep axis must be 0 or 1");
    if (n && !aabbs) return fail(c, CB2_ERR_ARG, "null argument");
    std::vector<cb2_aabb3> lifted(n);
    for (uint64_t i = 0; i < n; ++i) {
        lifted[i].min[0] = aabbs[i].min[0];
        lifted[i].min[1] = aabbs[i].min[1];
        lifted[i].min[2] = 0.0f;
        lifted[i].max[0] = aabbs[i].max[0];
        lifted[i].max[1] = aabbs[i].max[1];
        lifted[i].max[2] = 1.0f;
    }
    return cb2_sap3_find_collider_pairs(c, lifted.data(), n, axis_inout, perm_out, pairs_out, cap,
                                        n_pairs);
}

// BruteForce::find_collider_pairs (brute_force.rs:20-39): same pair SET as the sweep, reported in
// the caller's own indices, left ascending then right ascending.
int cb2_brute_force_find_collider_pairs_dev(cb2_ctx* c, const cb2_aabb3* aabbs, uint64_t n,
                                            uint32_t* pairs_out, uint64_t cap, uint64_t* n_pairs,
                                            void* stream) {
    if (!c || !n_pairs) return fail(c, CB2_ERR_ARG, "null argument");
    *n_pairs = 0;
    if (n == 0) return CB2_OK;
    if (!aabbs || (!pairs_out && cap)) return fail(c, CB2_ERR_ARG, "null argument");
    CK(c, cudaSetDevice(c->device));
    const char* err = "";
    int rc = sap_find_pairs(c->sap, aabbs, n, 0, nullptr, pairs_out, cap, n_pairs,
                            static_cast<cudaStream_t>(stream), &c->launches, &err, true);
    if (rc == CB2_ERR_NOSPC) c->err = "pairs_out too small; *n_pairs holds the required capacity";
    else if (rc) c->err = err;
    return rc;
}

int cb2_brute_force_find_collider_pairs(cb2_ctx* c, const cb2_aabb3* aabbs, uint64_t n,
                                        uint32_t* pairs_out, uint64_t cap, uint64_t* n_pairs) {
    if (!c || !n_pairs) return fail(c, CB2_ERR_ARG, "null argument");
    *n_pairs = 0;
    if (n == 0) return CB2_OK;
    if (!aabbs || (!pairs_out && cap)) return fail(c, CB2_ERR_ARG, "null argument");
    CK(c, cudaSetDevice(c->device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_box = 0, o_pairs = al(sizeof(cb2_aabb3) * n);
    uint64_t dcap = std::min<uint64_t>(cap, 16 * n + 1024);  // see cb2_sap3_find_collider_pairs
    int rc = CB2_OK;
    cudaStream_t st = c->stream[0];
    char* base = nullptr;
    for (int attempt = 0; attempt < 2; ++attempt) {
        rc = ensure_misc(c, o_pairs + al(8 * (dcap ? dcap : 1)));
        if (rc) return rc;
        base = static_cast<char*>(c->d_misc);
        CK(c, cudaMemcpyAsync(base + o_box, aabbs, sizeof(cb2_aabb3) * n, cudaMemcpyHostToDevice, st));
        rc = cb2_brute_force_find_collider_pairs_dev(c, reinterpret_cast<cb2_aabb3*>(base + o_box), n,
                                                     reinterpret_cast<uint32_t*>(base + o_pairs), dcap,
                                                     n_pairs, st);
        if (rc != CB2_ERR_NOSPC || *n_pairs > cap || dcap >= cap) break;
        dcap = *n_pairs;
    }
    if (rc) return rc;
    if (*n_pairs)
        CK(c, cudaMemcpyAsync(pairs_out, base + o_pairs, 8 * (*n_pairs), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    return CB2_OK;
}

int cb2_sap3_find_collider_pairs(cb2_ctx* c, const cb2_aabb3* aabbs, uint64_t n,
                                 uint32_t* axis_inout, uint32_t* perm_out, uint32_t* pairs_out,
                                 uint64_t cap, uint64_t* n_pairs) {
    if (!c || !axis_inout || !n_pairs) return fail(c, CB2_ERR_ARG, "null argument");
    *n_pairs = 0;
    if (*axis_inout > 2) return fail(c, CB2_ERR_ARG, "sweep axis must be 0, 1 or 2");
    if (n == 0) return CB2_OK;
    if (!aabbs || !perm_out || (!pairs_out && cap)) return fail(c, CB2_ERR_ARG, "null argument");
    if (n == 1) {  // sweep_prune.rs:83-85: early return, axis untouched
        perm_out[0] = 0;
        return CB2_OK;
    }
    CK(c, cudaSetDevice(c->device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_box = 0, o_perm = al(n * sizeof(cb2_aabb3)), o_pairs = o_perm + al(n * 4);
    // the device pair buffer is sized for what the call can need, not for the caller's bound (a
    // safe cap such as n(n-1)/2 must not turn into a multi-GB allocation): 16 pairs per box to
    // start with, the exact count on the one retry
    uint64_t dcap = std::min<uint64_t>(cap, 16 * n + 1024);
    int rc = CB2_OK;
    cudaStream_t st = c->stream[0];
    char* base = nullptr;
    for (int attempt = 0; attempt < 2; ++attempt) {
        rc = ensure_misc(c, o_pairs + al(dcap * 8));
        if (rc) return rc;
        base = static_cast<char*>(c->d_misc);
        CK(c, cudaMemcpyAsync(base + o_box, aabbs, n * sizeof(cb2_aabb3), cudaMemcpyHostToDevice, st));
        rc = cb2_sap3_find_collider_pairs_dev(c, reinterpret_cast<cb2_aabb3*>(base + o_box), n,
                                              *axis_inout, reinterpret_cast<uint32_t*>(base + o_perm),
                                              reinterpret_cast<uint32_t*>(base + o_pairs), dcap, n_pairs,
                                              st);
        if (rc != CB2_ERR_NOSPC || *n_pairs > cap || dcap >= cap) break;
        dcap = *n_pairs;  // fits the caller's buffer: once more with room for every pair
    }
    if (rc && rc != CB2_ERR_NOSPC) return rc;
    CK(c, cudaMemcpyAsync(perm_out, base + o_perm, n * 4, cudaMemcpyDeviceToHost, st));
    if (rc == CB2_OK && *n_pairs)
        CK(c, cudaMemcpyAsync(pairs_out, base + o_pairs, *n_pairs * 8, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    // compute_axis sees zero sums => axis 0 (sweep_prune.rs:130-133, 254-277); on NOSPC the
    // axis is left alone so the caller can retry the same call with a larger buffer
    if (rc == CB2_OK) *axis_inout = 0;
    return rc;
}

// ---- world AABBs ---------------------------------------------------------------------------------------
int cb2_world_aabbs_dev(cb2_ctx* c, const uint32_t* shape, const cb2_xform* t, uint64_t n,
                        cb2_aabb3* out, void* stream) {
    if (!c) return CB2_ERR_ARG;
    if (n == 0) return CB2_OK;
    if (!shape || !t || !out) return fail(c, CB2_ERR_ARG, "null argument");
    if (!c->d_shapes) return fail(c, CB2_ERR_ARG, "cb2_shapes_upload has not been called");
    CK(c, cudaSetDevice(c->device));
    aabb_args A;
    A.T.shapes = c->d_shapes;
    A.T.verts = c->d_verts;
    A.T.cells = c->d_cells;
    A.T.ext = c->d_cell_ext;
    A.shape = shape;
    A.t = reinterpret_cast<const float4*>(t);
    A.n = n;
    A.out = out;
    cudaError_t e = launch_world_aabbs(A, static_cast<cudaStream_t>(stream));
    ++c->launches;
    if (e != cudaSuccess) {
        c->err = std::string("kernel launch: ") + cudaGetErrorString(e);
        return CB2_ERR_CUDA;
    }
    return CB2_OK;
}

int cb2_world_aabbs(cb2_ctx* c, const uint32_t* shape, const cb2_xform* t, uint64_t n,
                    cb2_aabb3* out) {
    if (!c) return CB2_ERR_ARG;
    if (n == 0) return CB2_OK;
    if (!shape || !t || !out) return fail(c, CB2_ERR_ARG, "null argument");
    for (uint64_t i = 0; i < n; ++i)
        if (shape[i] >= c->n_shapes) return fail(c, CB2_ERR_ARG, "shape index out of range");
    CK(c, cudaSetDevice(c->device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_sh = 0, o_t = al(n * 4), o_out = o_t + al(n * sizeof(cb2_xform));
    int rc = ensure_misc(c, o_out + al(n * sizeof(cb2_aabb3)));
    if (rc) return rc;
    char* base = static_cast<char*>(c->d_misc);
    cudaStream_t st = c->stream[0];
    CK(c, cudaMemcpyAsync(base + o_sh, shape, n * 4, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(base + o_t, t, n * sizeof(cb2_xform), cudaMemcpyHostToDevice, st));
    rc = cb2_world_aabbs_dev(c, reinterpret_cast<uint32_t*>(base + o_sh),
                             reinterpret_cast<cb2_xform*>(base + o_t), n,
                             reinterpret_cast<cb2_aabb3*>(base + o_out), st);
    if (rc) return rc;
    CK(c, cudaMemcpyAsync(out, base + o_out, n * sizeof(cb2_aabb3), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    return CB2_OK;
}

}  // extern "C"

// ---- composite shapes (gjk/mod.rs:412-588) ------------------------------------------------------------
extern "C" int cb2_composites_upload(cb2_ctx* c, const uint32_t* part_off, const uint32_t* part_shape,
                                     const cb2_xform* part_local, uint32_t n_comp) {
    if (!c || !part_off || (n_comp && part_off[n_comp] && (!part_shape || !part_local)))
        return fail(c, CB2_ERR_ARG, "null argument");
    if (!c->d_shapes) return fail(c, CB2_ERR_ARG, "cb2_shapes_upload has not been called");
    const uint32_t n_parts = part_off[n_comp];
    if (part_off[0] != 0) return fail(c, CB2_ERR_ARG, "part_off[0] must be 0");
    for (uint32_t k = 0; k < n_comp; ++k)
        if (part_off[k] > part_off[k + 1]) return fail(c, CB2_ERR_ARG, "part_off must not decrease");
    for (uint32_t k = 0; k < n_parts; ++k)
        if (part_shape[k] >= c->n_shapes) return fail(c, CB2_ERR_ARG, "part shape index out of range");
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaDeviceSynchronize());
    if (c->d_comp_off) cudaFree(c->d_comp_off);
    if (c->d_comp_shape) cudaFree(c->d_comp_shape);
    if (c->d_comp_local) cudaFree(c->d_comp_local);
    c->d_comp_off = c->d_comp_shape = nullptr;
    c->d_comp_local = nullptr;
    CK(c, cudaMalloc(&c->d_comp_off, sizeof(uint32_t) * ((size_t)n_comp + 1)));
    CK(c, cudaMalloc(&c->d_comp_shape, sizeof(uint32_t) * (n_parts ? n_parts : 1)));
    CK(c, cudaMalloc(&c->d_comp_local, sizeof(cb2_xform) * (n_parts ? n_parts : 1)));
    CK(c, cudaMemcpy(c->d_comp_off, part_off, sizeof(uint32_t) * ((size_t)n_comp + 1), cudaMemcpyHostToDevice));
    if (n_parts) {
        CK(c, cudaMemcpy(c->d_comp_shape, part_shape, sizeof(uint32_t) * n_parts, cudaMemcpyHostToDevice));
        CK(c, cudaMemcpy(c->d_comp_local, part_local, sizeof(cb2_xform) * n_parts, cudaMemcpyHostToDevice));
    }
    c->n_comp = n_comp;
    c->h_comp_off.assign(part_off, part_off + n_comp + 1);
    return CB2_OK;
}

static int narrow2_dev(cb2_ctx* c, int op, uint32_t strategy, const cb2_settings* settings,
                       const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform2* l_t0,
                       const cb2_xform2* l_t1, const cb2_xform2* r_t0, const cb2_xform2* r_t1, uint64_t n,
                       uint32_t iter_cap, cb2_contact2* out_c, float* out_d, uint8_t* status,
                       cudaStream_t st, int slot, const uint32_t* pair_body = nullptr);

// d2: the 2-D instantiation — cb2_xform2 transforms (same 32-byte stride as cb2_xform), cb2_contact2
// records, the 2-D composite table and the GJK2 kernels
static int complex_host(cb2_ctx* c, bool d2, op_kind op, uint32_t strategy, const cb2_settings* settings,
                        const uint32_t* l_comp, const uint32_t* r_comp, const void* l_t0,
                        const void* l_t1, const void* r_t0, const void* r_t1,
                        uint64_t n, uint32_t iter_cap, void* out_c, float* out_d,
                        uint8_t* status) {
    if (!c) return CB2_ERR_ARG;
    static_assert(sizeof(cb2_xform) == sizeof(cb2_xform2), "both transforms travel as 2 x float4");
    const uint32_t* d_off = d2 ? c->d_comp2_off : c->d_comp_off;
    const uint32_t* d_cshape = d2 ? c->d_comp2_shape : c->d_comp_shape;
    const float4* d_clocal = d2 ? c->d_comp2_local : c->d_comp_local;
    const uint32_t n_comp = d2 ? c->n_comp2 : c->n_comp;
    const std::vector<uint32_t>& h_off = d2 ? c->h_comp2_off : c->h_comp_off;
    const size_t contact_bytes = d2 ? sizeof(cb2_contact2) : sizeof(cb2_contact);
    int rc = check_settings(c, settings, d2);
    if (rc) return rc;
    if (op != OP_DISTANCE && strategy > CB2_COLLISION_ONLY) return fail(c, CB2_ERR_ARG, "bad strategy");
    if (n == 0) return CB2_OK;
    if (!d_off) return fail(c, CB2_ERR_ARG, "the composite table has not been uploaded");
    if (!l_comp || !r_comp || !l_t0 || !r_t0 || !status) return fail(c, CB2_ERR_ARG, "null argument");
    if (op == OP_DISTANCE ? !out_d : !out_c) return fail(c, CB2_ERR_ARG, "null output");
    if (op == OP_TOI && (!l_t1 || !r_t1)) return fail(c, CB2_ERR_ARG, "null end transform");
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (l_comp[i] >= n_comp || r_comp[i] >= n_comp)
            return fail(c, CB2_ERR_ARG, "composite index out of range");
        total += (uint64_t)(h_off[l_comp[i] + 1] - h_off[l_comp[i]]) *
                 (h_off[r_comp[i] + 1] - h_off[r_comp[i]]);
    }
    if (total > 0x7FFFFFFFull) return fail(c, CB2_ERR_ARG, "more than 2^31-1 part pairs in one call");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream[0];
    const int n_xf = op == OP_TOI ? 4 : 2;
    const size_t out_rec = op == OP_DISTANCE ? sizeof(float) : contact_bytes;
    const uint64_t T = total ? total : 1;
    size_t scan_tmp = 0;
    cx_scan(nullptr, &scan_tmp, nullptr, nullptr, n, st);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += al(bytes); return o; };
    const size_t o_lc = take(4 * n), o_rc = take(4 * n);
    size_t o_t[4];
    for (int x = 0; x < n_xf; ++x) o_t[x] = take(sizeof(cb2_xform) * n);
    const size_t o_cnt = take(8 * n), o_off = take(8 * n), o_scan = take(scan_tmp);
    const size_t o_sl = take(4 * T), o_sr = take(4 * T);
    size_t o_st[4];
    for (int x = 0; x < n_xf; ++x) o_st[x] = take(sizeof(cb2_xform) * T);
    const size_t o_sout = take(out_rec * T), o_sstat = take(T);
    const size_t o_out = take(out_rec * n), o_stat = take(n);
    if (c->cx_cap < off) {
        if (c->d_cx) cudaFree(c->d_cx);
        c->d_cx = nullptr;
        c->cx_cap = 0;
        CK(c, cudaMalloc(&c->d_cx, off));
        c->cx_cap = off;
    }
    char* base = static_cast<char*>(c->d_cx);
    // order of the transform slots: l_t0, r_t0, l_t1, r_t1
    const void* src[4] = {l_t0, r_t0, l_t1, r_t1};
    CK(c, cudaMemcpyAsync(base + o_lc, l_comp, 4 * n, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(base + o_rc, r_comp, 4 * n, cudaMemcpyHostToDevice, st));
    for (int x = 0; x < n_xf; ++x)
        CK(c, cudaMemcpyAsync(base + o_t[x], src[x], sizeof(cb2_xform) * n, cudaMemcpyHostToDevice, st));
    complex_args X;
    std::memset(&X, 0, sizeof(X));
    X.dim2 = d2;
    X.comp_off = d_off;
    X.comp_shape = d_cshape;
    X.comp_local = d_clocal;
    X.l_comp = reinterpret_cast<uint32_t*>(base + o_lc);
    X.r_comp = reinterpret_cast<uint32_t*>(base + o_rc);
    X.l_t0 = reinterpret_cast<float4*>(base + o_t[0]);
    X.r_t0 = reinterpret_cast<float4*>(base + o_t[1]);
    if (op == OP_TOI) {
        X.l_t1 = reinterpret_cast<float4*>(base + o_t[2]);
        X.r_t1 = reinterpret_cast<float4*>(base + o_t[3]);
    }
    X.n = n;
    X.sub_count = reinterpret_cast<uint64_t*>(base + o_cnt);
    X.sub_off = reinterpret_cast<uint64_t*>(base + o_off);
    X.sub_l_shape = reinterpret_cast<uint32_t*>(base + o_sl);
    X.sub_r_shape = reinterpret_cast<uint32_t*>(base + o_sr);
    X.sub_l_t0 = reinterpret_cast<float4*>(base + o_st[0]);
    X.sub_r_t0 = reinterpret_cast<float4*>(base + o_st[1]);
    if (op == OP_TOI) {
        X.sub_l_t1 = reinterpret_cast<float4*>(base + o_st[2]);
        X.sub_r_t1 = reinterpret_cast<float4*>(base + o_st[3]);
    }
    X.sub_contact = base + o_sout;
    X.sub_distance = reinterpret_cast<float*>(base + o_sout);
    X.sub_status = reinterpret_cast<uint8_t*>(base + o_sstat);
    X.out_contact = base + o_out;
    X.out_distance = reinterpret_cast<float*>(base + o_out);
    X.out_status = reinterpret_cast<uint8_t*>(base + o_stat);
    CK(c, launch_cx_count(X, st));
    size_t tb = scan_tmp;
    CK(c, cx_scan(base + o_scan, &tb, X.sub_count, X.sub_off, n, st));
    CK(c, launch_cx_expand(X, st));
    c->launches += 3;
    if (total && d2) {
        rc = narrow2_dev(c, op == OP_INTERSECTION ? 0 : (op == OP_DISTANCE ? 1 : 2),
                         op == OP_INTERSECTION ? strategy : CB2_FULL_RESOLUTION, settings, X.sub_l_shape,
                         X.sub_r_shape, reinterpret_cast<const cb2_xform2*>(X.sub_l_t0),
                         reinterpret_cast<const cb2_xform2*>(X.sub_l_t1),
                         reinterpret_cast<const cb2_xform2*>(X.sub_r_t0),
                         reinterpret_cast<const cb2_xform2*>(X.sub_r_t1), total, iter_cap,
                         reinterpret_cast<cb2_contact2*>(base + o_sout), reinterpret_cast<float*>(base + o_sout),
                         reinterpret_cast<uint8_t*>(base + o_sstat), st, 0);
        if (rc) return rc;
    } else if (total) {
        narrow_args A;
        std::memset(&A, 0, sizeof(A));
        A.l_shape = X.sub_l_shape;
        A.r_shape = X.sub_r_shape;
        A.l_t = X.sub_l_t0;
        A.r_t = X.sub_r_t0;
        A.l_t1 = X.sub_l_t1;
        A.r_t1 = X.sub_r_t1;
        A.n = total;
        A.settings = *settings;
        A.iter_cap = iter_cap ? iter_cap : 1024u;
        A.out_contact = reinterpret_cast<cb2_contact*>(base + o_sout);
        A.out_distance = reinterpret_cast<float*>(base + o_sout);
        A.out_status = reinterpret_cast<uint8_t*>(base + o_sstat);
        rc = launch_op(c, op, op == OP_INTERSECTION ? strategy : CB2_FULL_RESOLUTION, A, st, 0);
        if (rc) return rc;
    }
    CK(c, launch_cx_reduce(X, op == OP_INTERSECTION ? 0u : (op == OP_DISTANCE ? 1u : 2u), strategy, st));
    ++c->launches;
    if (op == OP_DISTANCE)
        CK(c, cudaMemcpyAsync(out_d, base + o_out, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    else
        CK(c, cudaMemcpyAsync(out_c, base + o_out, contact_bytes * n, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(status, base + o_stat, n, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    return CB2_OK;
}

extern "C" {
int cb2_gjk3_intersection_complex(cb2_ctx* c, uint32_t strategy, const cb2_settings* settings,
                                  const uint32_t* l_comp, const uint32_t* r_comp,
                                  const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                                  cb2_contact* out, uint8_t* status) {
    return complex_host(c, false, OP_INTERSECTION, strategy, settings, l_comp, r_comp, l_t, nullptr, r_t,
                        nullptr, n, 0, out, nullptr, status);
}
int cb2_gjk3_distance_complex(cb2_ctx* c, const cb2_settings* settings, const uint32_t* l_comp,
                              const uint32_t* r_comp, const cb2_xform* l_t, const cb2_xform* r_t,
                              uint64_t n, float* out, uint8_t* status) {
    return complex_host(c, false, OP_DISTANCE, 0, settings, l_comp, r_comp, l_t, nullptr, r_t, nullptr, n, 0,
                        nullptr, out, status);
}
int cb2_gjk3_time_of_impact_complex(cb2_ctx* c, uint32_t strategy, const cb2_settings* settings,
                                    const uint32_t* l_comp, const uint32_t* r_comp,
                                    const cb2_xform* l_t0, const cb2_xform* l_t1,
                                    const cb2_xform* r_t0, const cb2_xform* r_t1, uint64_t n,
                                    uint32_t iter_cap, cb2_contact* out, uint8_t* status) {
    return complex_host(c, false, OP_TOI, strategy, settings, l_comp, r_comp, l_t0, l_t1, r_t0, r_t1, n,
                        iter_cap, out, nullptr, status);
}
}

// ==== 2-D narrow phase: GJK2 / EPA2 (cb2_narrow2d.cu) ================================================
static int narrow2_check(cb2_ctx* c, int op, uint32_t strategy, const cb2_settings* settings,
                         const void* l_shape, const void* r_shape, const void* l_t0, const void* l_t1,
                         const void* r_t0, const void* r_t1, const void* out, const void* status) {
    if (!c) return CB2_ERR_ARG;
    int rc = check_settings(c, settings, true);
    if (rc) return rc;
    if (op == 0 && strategy > CB2_COLLISION_ONLY) return fail(c, CB2_ERR_ARG, "bad strategy");
    if (!c->d_shapes2) return fail(c, CB2_ERR_ARG, "cb2_shapes2_upload has not been called");
    if (!l_shape || !r_shape || !l_t0 || !r_t0 || !status || !out) return fail(c, CB2_ERR_ARG, "null argument");
    if (op == 2 && (!l_t1 || !r_t1)) return fail(c, CB2_ERR_ARG, "null end transform");
    return CB2_OK;
}

// GJK2 -> EPA2 hand-off buffers of one compute slot (hit list + two queues, 3-point simplexes)
static int ensure_ws2(cb2_ctx* c, int slot, uint64_t n) {
    if (c->ws2_cap[slot] >= n) return CB2_OK;
    if (c->d_hits2[slot]) cudaFree(c->d_hits2[slot]);
    if (c->d_sx2[slot]) cudaFree(c->d_sx2[slot]);
    c->d_hits2[slot] = nullptr;
    c->d_sx2[slot] = nullptr;
    c->ws2_cap[slot] = 0;
    CK(c, cudaMalloc(&c->d_hits2[slot], 3 * n * sizeof(uint32_t)));
    CK(c, cudaMalloc(&c->d_sx2[slot], n * 48));
    c->ws2_cap[slot] = n;
    return CB2_OK;
}

static int narrow2_dev(cb2_ctx* c, int op, uint32_t strategy, const cb2_settings* settings,
                       const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform2* l_t0,
                       const cb2_xform2* l_t1, const cb2_xform2* r_t0, const cb2_xform2* r_t1, uint64_t n,
                       uint32_t iter_cap, cb2_contact2* out_c, float* out_d, uint8_t* status,
                       cudaStream_t st, int slot, const uint32_t* pair_body) {
    narrow2_args A;
    std::memset(&A, 0, sizeof(A));
    A.pair_body = pair_body;
    const bool epa = op == 0 && strategy == CB2_FULL_RESOLUTION;
    if (!c->d_cnt2[slot]) CK(c, cudaMalloc(&c->d_cnt2[slot], 8 * sizeof(uint32_t)));
    A.counters = c->d_cnt2[slot];  // queue cursors of the persistent kernels
    A.long_tails = c->long_tails2;
    if (epa) {
        int rc2 = ensure_ws2(c, slot, n);
        if (rc2) return rc2;
        A.counters = c->d_cnt2[slot];
        A.hits = c->d_hits2[slot];
        A.simplex = c->d_sx2[slot];
    }
    A.shapes = c->d_shapes2;
    A.verts = c->d_verts2;
    A.l_shape = l_shape;
    A.r_shape = r_shape;
    A.l_t = l_t0;
    A.r_t = r_t0;
    A.l_t1 = l_t1;
    A.r_t1 = r_t1;
    A.n = n;
    A.settings = *settings;
    A.iter_cap = iter_cap ? iter_cap : 1024u;
    A.strategy = strategy;
    A.out_contact = out_c;
    A.out_distance = out_d;
    A.out_status = status;
    CK(c, launch_narrow2(op, A, c->sm_count, st));
    c->launches += epa ? 3 : 1;
    return CB2_OK;
}

// Host-buffer path: the same pipeline as narrow_host.
static int narrow2_host(cb2_ctx* c, int op, uint32_t strategy, const cb2_settings* settings,
                        const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform2* l_t0,
                        const cb2_xform2* l_t1, const cb2_xform2* r_t0, const cb2_xform2* r_t1,
                        uint64_t n, uint32_t iter_cap, cb2_contact2* out_c, float* out_d,
                        uint8_t* status) {
    int rc = narrow2_check(c, op, strategy, settings, l_shape, r_shape, l_t0, l_t1, r_t0, r_t1,
                           op == 1 ? (const void*)out_d : (const void*)out_c, status);
    if (rc) return rc;
    if (n == 0) return CB2_OK;
    CK(c, cudaSetDevice(c->device));
    const uint64_t chunk = n < c->chunk_pairs ? n : c->chunk_pairs;
    const int n_xf = op == 2 ? 4 : 2;
    const size_t out_rec = op == 1 ? sizeof(float) : sizeof(cb2_contact2);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t xf_bytes = al(chunk * sizeof(cb2_xform2));
    const size_t o_ls = n_xf * xf_bytes;
    const size_t o_rs = o_ls + al(chunk * 4);
    const size_t o_out = o_rs + al(chunk * 4);
    const size_t o_st = o_out + al(chunk * out_rec);
    rc = pipeline_prepare(c, o_st + al(chunk));
    if (rc) return rc;
    if (op == 0 && strategy == CB2_FULL_RESOLUTION)
        for (int slot = 0; slot < c->n_cstreams; ++slot) {
            rc = ensure_ws2(c, slot, chunk);
            if (rc) return rc;
        }
    const cb2_xform2* xf_src[4] = {l_t0, r_t0, l_t1, r_t1};
    const size_t xs = xf_bytes / sizeof(cb2_xform2);
    return pipeline_run(
        c, n,
        [&](uint64_t b, uint64_t m, char* base, cudaStream_t st) -> int {
            if (!indices_in_range(l_shape + b, r_shape + b, m, c->n_shapes2))
                return fail(c, CB2_ERR_ARG, "shape index out of range");
            for (int x = 0; x < n_xf; ++x)
                CK(c, cudaMemcpyAsync(base + x * xf_bytes, xf_src[x] + b, m * sizeof(cb2_xform2),
                                      cudaMemcpyHostToDevice, st));
            CK(c, cudaMemcpyAsync(base + o_ls, l_shape + b, m * 4, cudaMemcpyHostToDevice, st));
            CK(c, cudaMemcpyAsync(base + o_rs, r_shape + b, m * 4, cudaMemcpyHostToDevice, st));
            return CB2_OK;
        },
        [&](uint64_t, uint64_t m, char* base, cudaStream_t st, int slot, cudaStream_t) -> int {
            const cb2_xform2* X = reinterpret_cast<const cb2_xform2*>(base);
            return narrow2_dev(c, op, strategy, settings, reinterpret_cast<const uint32_t*>(base + o_ls),
                               reinterpret_cast<const uint32_t*>(base + o_rs), X, X + 2 * xs, X + xs,
                               X + 3 * xs, m, iter_cap, reinterpret_cast<cb2_contact2*>(base + o_out),
                               reinterpret_cast<float*>(base + o_out),
                               reinterpret_cast<uint8_t*>(base + o_st), st, slot);
        },
        [&](uint64_t b, uint64_t m, char* base, cudaStream_t st) -> int {
            if (op == 1)
                CK(c, cudaMemcpyAsync(out_d + b, base + o_out, m * sizeof(float), cudaMemcpyDeviceToHost, st));
            else
                CK(c, cudaMemcpyAsync(out_c + b, base + o_out, m * sizeof(cb2_contact2),
                                      cudaMemcpyDeviceToHost, st));
            CK(c, cudaMemcpyAsync(status + b, base + o_st, m, cudaMemcpyDeviceToHost, st));
            return CB2_OK;
        });
}

extern "C" {

int cb2_shapes2_upload(cb2_ctx* c, const cb2_shape2* shapes, uint32_t n_shapes, const float* verts_xy,
                       uint64_t n_verts) {
    if (!c) return CB2_ERR_ARG;
    if (!shapes || n_shapes == 0) return fail(c, CB2_ERR_ARG, "empty shape table");
    bool has_circle = false, has_other = false, has_walk = false;
    for (uint32_t i = 0; i < n_shapes; ++i) {
        const cb2_shape2& s = shapes[i];
        if (s.kind > CB2_POLYGON) return fail(c, CB2_ERR_ARG, "unknown 2-D shape kind");
        has_circle |= s.kind == CB2_CIRCLE;
        has_other |= s.kind != CB2_CIRCLE;
        has_walk |= s.kind == CB2_POLYGON && s.vtx_cnt >= 10;
        if (s.kind == CB2_POLYGON) {
            if ((uint64_t)s.vtx_off + s.vtx_cnt > n_verts || (s.vtx_cnt && !verts_xy))
                return fail(c, CB2_ERR_ARG, "polygon vertex range out of bounds");
        }
    }
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaDeviceSynchronize());
    drop_composites(c, true);
    if (c->d_shapes2) cudaFree(c->d_shapes2);
    if (c->d_verts2) cudaFree(c->d_verts2);
    c->d_shapes2 = nullptr;
    c->d_verts2 = nullptr;
    c->n_shapes2 = 0;
    CK(c, cudaMalloc(&c->d_shapes2, sizeof(cb2_shape2) * n_shapes));
    CK(c, cudaMemcpy(c->d_shapes2, shapes, sizeof(cb2_shape2) * n_shapes, cudaMemcpyHostToDevice));
    // at least one (unused) vertex so polygon-free tables still have a valid pointer
    CK(c, cudaMalloc(&c->d_verts2, sizeof(float) * 2 * (n_verts ? n_verts : 1)));
    if (n_verts)
        CK(c, cudaMemcpy(c->d_verts2, verts_xy, sizeof(float) * 2 * n_verts, cudaMemcpyHostToDevice));
    c->n_shapes2 = n_shapes;
    c->long_tails2 = (has_circle && has_other) || has_walk;
    return CB2_OK;
}

int cb2_gjk2_intersection(cb2_ctx* c, uint32_t strategy, const cb2_settings* settings,
                          const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform2* l_t,
                          const cb2_xform2* r_t, uint64_t n, cb2_contact2* out, uint8_t* status) {
    return narrow2_host(c, 0, strategy, settings, l_shape, r_shape, l_t, nullptr, r_t, nullptr, n, 0, out,
                        nullptr, status);
}
int cb2_gjk2_distance(cb2_ctx* c, const cb2_settings* settings, const uint32_t* l_shape,
                      const uint32_t* r_shape, const cb2_xform2* l_t, const cb2_xform2* r_t, uint64_t n,
                      float* out, uint8_t* status) {
    return narrow2_host(c, 1, 0, settings, l_shape, r_shape, l_t, nullptr, r_t, nullptr, n, 0, nullptr, out,
                        status);
}
int cb2_gjk2_time_of_impact(cb2_ctx* c, const cb2_settings* settings, const uint32_t* l_shape,
                            const uint32_t* r_shape, const cb2_xform2* l_t0, const cb2_xform2* l_t1,
                            const cb2_xform2* r_t0, const cb2_xform2* r_t1, uint64_t n, uint32_t iter_cap,
                            cb2_contact2* out, uint8_t* status) {
    return narrow2_host(c, 2, 0, settings, l_shape, r_shape, l_t0, l_t1, r_t0, r_t1, n, iter_cap, out,
                        nullptr, status);
}
int cb2_gjk2_intersection_dev(cb2_ctx* c, uint32_t strategy, const cb2_settings* settings,
                              const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform2* l_t,
                              const cb2_xform2* r_t, uint64_t n, cb2_contact2* out, uint8_t* status,
                              void* stream) {
    int rc = narrow2_check(c, 0, strategy, settings, l_shape, r_shape, l_t, nullptr, r_t, nullptr, out, status);
    if (rc || n == 0) return rc;
    CK(c, cudaSetDevice(c->device));
    return narrow2_dev(c, 0, strategy, settings, l_shape, r_shape, l_t, nullptr, r_t, nullptr, n, 0, out,
                       nullptr, status, static_cast<cudaStream_t>(stream), 0);
}
int cb2_gjk2_distance_dev(cb2_ctx* c, const cb2_settings* settings, const uint32_t* l_shape,
                          const uint32_t* r_shape, const cb2_xform2* l_t, const cb2_xform2* r_t,
                          uint64_t n, float* out, uint8_t* status, void* stream) {
    int rc = narrow2_check(c, 1, 0, settings, l_shape, r_shape, l_t, nullptr, r_t, nullptr, out, status);
    if (rc || n == 0) return rc;
    CK(c, cudaSetDevice(c->device));
    return narrow2_dev(c, 1, 0, settings, l_shape, r_shape, l_t, nullptr, r_t, nullptr, n, 0, nullptr, out,
                       status, static_cast<cudaStream_t>(stream), 0);
}
int cb2_gjk2_time_of_impact_dev(cb2_ctx* c, const cb2_settings* settings, const uint32_t* l_shape,
                                const uint32_t* r_shape, const cb2_xform2* l_t0, const cb2_xform2* l_t1,
                                const cb2_xform2* r_t0, const cb2_xform2* r_t1, uint64_t n,
                                uint32_t iter_cap, cb2_contact2* out, uint8_t* status, void* stream) {
    int rc = narrow2_check(c, 2, 0, settings, l_shape, r_shape, l_t0, l_t1, r_t0, r_t1, out, status);
    if (rc || n == 0) return rc;
    CK(c, cudaSetDevice(c->device));
    return narrow2_dev(c, 2, 0, settings, l_shape, r_shape, l_t0, l_t1, r_t0, r_t1, n, iter_cap, out,
                       nullptr, status, static_cast<cudaStream_t>(stream), 0);
}

}  // extern "C"

// ---- 2-D composite shapes: GJK2::intersection_complex / distance_complex / ..._time_of_impact ----------
extern "C" {
int cb2_composites2_upload(cb2_ctx* c, const uint32_t* part_off, const uint32_t* part_shape,
                           const cb2_xform2* part_local, uint32_t n_comp) {
    if (!c || !part_off || (n_comp && part_off[n_comp] && (!part_shape || !part_local)))
        return fail(c, CB2_ERR_ARG, "null argument");
    if (!c->d_shapes2) return fail(c, CB2_ERR_ARG, "cb2_shapes2_upload has not been called");
    const uint32_t n_parts = part_off[n_comp];
    if (part_off[0] != 0) return fail(c, CB2_ERR_ARG, "part_off[0] must be 0");
    for (uint32_t k = 0; k < n_comp; ++k)
        if (part_off[k] > part_off[k + 1]) return fail(c, CB2_ERR_ARG, "part_off must not decrease");
    for (uint32_t k = 0; k < n_parts; ++k)
        if (part_shape[k] >= c->n_shapes2) return fail(c, CB2_ERR_ARG, "part shape index out of range");
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaDeviceSynchronize());
    if (c->d_comp2_off) cudaFree(c->d_comp2_off);
    if (c->d_comp2_shape) cudaFree(c->d_comp2_shape);
    if (c->d_comp2_local) cudaFree(c->d_comp2_local);
    c->d_comp2_off = c->d_comp2_shape = nullptr;
    c->d_comp2_local = nullptr;
    c->n_comp2 = 0;
    CK(c, cudaMalloc(&c->d_comp2_off, sizeof(uint32_t) * ((size_t)n_comp + 1)));
    CK(c, cudaMalloc(&c->d_comp2_shape, sizeof(uint32_t) * (n_parts ? n_parts : 1)));
    CK(c, cudaMalloc(&c->d_comp2_local, sizeof(cb2_xform2) * (n_parts ? n_parts : 1)));
    CK(c, cudaMemcpy(c->d_comp2_off, part_off, sizeof(uint32_t) * ((size_t)n_comp + 1), cudaMemcpyHostToDevice));
    if (n_parts) {
        CK(c, cudaMemcpy(c->d_comp2_shape, part_shape, sizeof(uint32_t) * n_parts, cudaMemcpyHostToDevice));
        CK(c, cudaMemcpy(c->d_comp2_local, part_local, sizeof(cb2_xform2) * n_parts, cudaMemcpyHostToDevice));
    }
    c->h_comp2_off.assign(part_off, part_off + n_comp + 1);
    c->n_comp2 = n_comp;
    return CB2_OK;
}
int cb2_gjk2_intersection_complex(cb2_ctx* c, uint32_t strategy, const cb2_settings* settings,
                                  const uint32_t* l_comp, const uint32_t* r_comp, const cb2_xform2* l_t,
                                  const cb2_xform2* r_t, uint64_t n, cb2_contact2* out, uint8_t* status) {
    return complex_host(c, true, OP_INTERSECTION, strategy, settings, l_comp, r_comp, l_t, nullptr, r_t, nullptr,
                        n, 0, out, nullptr, status);
}
int cb2_gjk2_distance_complex(cb2_ctx* c, const cb2_settings* settings, const uint32_t* l_comp,
                              const uint32_t* r_comp, const cb2_xform2* l_t, const cb2_xform2* r_t,
                              uint64_t n, float* out, uint8_t* status) {
    return complex_host(c, true, OP_DISTANCE, 0, settings, l_comp, r_comp, l_t, nullptr, r_t, nullptr, n, 0,
                        nullptr, out, status);
}
int cb2_gjk2_time_of_impact_complex(cb2_ctx* c, uint32_t strategy, const cb2_settings* settings,
                                    const uint32_t* l_comp, const uint32_t* r_comp, const cb2_xform2* l_t0,
                                    const cb2_xform2* l_t1, const cb2_xform2* r_t0, const cb2_xform2* r_t1,
                                    uint64_t n, uint32_t iter_cap, cb2_contact2* out, uint8_t* status) {
    return complex_host(c, true, OP_TOI, strategy, settings, l_comp, r_comp, l_t0, l_t1, r_t0, r_t1, n, iter_cap,
                        out, nullptr, status);
}
}  // extern "C"

// ---- 2-D world bounds (the step before SweepAndPrune2) ---------------------------------------------------
extern "C" {
int cb2_world_aabbs2_dev(cb2_ctx* c, const uint32_t* shape, const cb2_xform2* t, uint64_t n, cb2_aabb2* out,
                         void* stream) {
    if (!c) return CB2_ERR_ARG;
    if (n == 0) return CB2_OK;
    if (!shape || !t || !out) return fail(c, CB2_ERR_ARG, "null argument");
    if (!c->d_shapes2) return fail(c, CB2_ERR_ARG, "cb2_shapes2_upload has not been called");
    CK(c, cudaSetDevice(c->device));
    CK(c, launch_world_aabbs2(c->d_shapes2, c->d_verts2, shape, t, n, out, static_cast<cudaStream_t>(stream)));
    ++c->launches;
    return CB2_OK;
}
int cb2_world_aabbs2(cb2_ctx* c, const uint32_t* shape, const cb2_xform2* t, uint64_t n, cb2_aabb2* out) {
    if (!c) return CB2_ERR_ARG;
    if (n == 0) return CB2_OK;
    if (!shape || !t || !out) return fail(c, CB2_ERR_ARG, "null argument");
    if (!c->d_shapes2) return fail(c, CB2_ERR_ARG, "cb2_shapes2_upload has not been called");
    for (uint64_t i = 0; i < n; ++i)
        if (shape[i] >= c->n_shapes2) return fail(c, CB2_ERR_ARG, "shape index out of range");
    CK(c, cudaSetDevice(c->device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_sh = 0, o_t = al(4 * n), o_out = o_t + al(sizeof(cb2_xform2) * n);
    int rc = ensure_misc(c, o_out + al(sizeof(cb2_aabb2) * n));
    if (rc) return rc;
    char* base = static_cast<char*>(c->d_misc);
    cudaStream_t st = c->stream[0];
    CK(c, cudaMemcpyAsync(base + o_sh, shape, 4 * n, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(base + o_t, t, sizeof(cb2_xform2) * n, cudaMemcpyHostToDevice, st));
    rc = cb2_world_aabbs2_dev(c, reinterpret_cast<uint32_t*>(base + o_sh), reinterpret_cast<cb2_xform2*>(base + o_t),
                              n, reinterpret_cast<cb2_aabb2*>(base + o_out), st);
    if (rc) return rc;
    CK(c, cudaMemcpyAsync(out, base + o_out, sizeof(cb2_aabb2) * n, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    return CB2_OK;
}
}  // extern "C"

// ---- the 2-D chain on the device: boxes -> SweepAndPrune2 pairs -> GJK2 on body pairs -------------------
extern "C" {
int cb2_sap2_find_collider_pairs_dev(cb2_ctx* c, const cb2_aabb2* aabbs, uint64_t n, uint32_t axis,
                                     uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap, uint64_t* n_pairs,
                                     void* stream) {
    if (!c || !n_pairs) return fail(c, CB2_ERR_ARG, "null argument");
    *n_pairs = 0;
    if (axis > 1) return fail(c, CB2_ERR_ARG, "sweep axis must be 0 or 1");
    if (n == 0) return CB2_OK;
    if (!aabbs || !perm_out || (!pairs_out && cap)) return fail(c, CB2_ERR_ARG, "null argument");
    CK(c, cudaSetDevice(c->device));
    if (c->lift_cap < n) {
        if (c->d_lift) cudaFree(c->d_lift);
        c->d_lift = nullptr;
        c->lift_cap = 0;
        CK(c, cudaMalloc(&c->d_lift, sizeof(cb2_aabb3) * n));
        c->lift_cap = n;
    }
    CK(c, launch_lift_aabb2(aabbs, n, c->d_lift, static_cast<cudaStream_t>(stream)));
    ++c->launches;
    return cb2_sap3_find_collider_pairs_dev(c, c->d_lift, n, axis, perm_out, pairs_out, cap, n_pairs, stream);
}
int cb2_gjk2_intersection_bodies_dev(cb2_ctx* c, uint32_t strategy, const cb2_settings* settings,
                                     const uint32_t* body_shape, const cb2_xform2* body_t,
                                     const uint32_t* pair_body, uint64_t n, cb2_contact2* out, uint8_t* status,
                                     void* stream) {
    int rc = narrow2_check(c, 0, strategy, settings, body_shape, body_shape, body_t, nullptr, body_t, nullptr, out,
                           status);
    if (rc || n == 0) return rc;
    if (!pair_body) return fail(c, CB2_ERR_ARG, "null argument");
    CK(c, cudaSetDevice(c->device));
    return narrow2_dev(c, 0, strategy, settings, body_shape, body_shape, body_t, nullptr, body_t, nullptr, n, 0, out,
                       nullptr, status, static_cast<cudaStream_t>(stream), 0, pair_body);
}
}  // extern "C"
