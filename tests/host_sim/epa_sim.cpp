// epa_sim.cpp — TEST INFRASTRUCTURE: runs the lane-per-pair EPA tier (csrc/cb2_epa_lane.cuh) and the
// thread-per-pair GJK (csrc/cb2_pair.cuh) on the CPU, compiled from the very same source as the
// CUDA kernels through csrc/cb2_hostsim.h, one lane at a time.  tests/test_host_sim.py compares the
// results bit for bit with the oracle, so the kernel's control flow (work-list fetch, the
// visibility mask walk, the edge list, dump / resume between tiers) is checked without a GPU.
// Never linked into the product library; the library itself has no CPU path.
//
// Shapes: every Primitive3 kind except the HalfEdge polyhedron (its ring tables are built by the
// upload path, which is CUDA-side).  Polyhedra run without support cells (cfg N = 0: full scan) —
// the cells are an exact accelerator checked separately on the GPU.
#define CB2_HOST_SIM 1
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../collision_rs_b200/csrc/cb2_epa_team.cuh"

using namespace cb2;

namespace {

struct sim_tables {
    std::vector<dshape> shapes;
    std::vector<float4> verts;
};

int build_tables(const cb2_shape* shapes, uint32_t n_shapes, const float* verts_xyz, uint64_t n_verts,
                 sim_tables& T) {
    T.shapes.resize(n_shapes);
    for (uint32_t i = 0; i < n_shapes; ++i) {
        const cb2_shape& s = shapes[i];
        dshape d;
        std::memset(&d, 0, sizeof(d));
        d.kind = s.kind;
        switch (s.kind) {
            case CB2_SPHERE: d.hx = s.p[0]; break;
            case CB2_CUBOID: d.hx = s.p[0] / 2.0f; d.hy = s.p[1] / 2.0f; d.hz = s.p[2] / 2.0f; break;
            case CB2_CUBE: d.kind = CB2_CUBOID; d.hx = d.hy = d.hz = s.p[0] / 2.0f; break;
            case CB2_QUAD: d.hx = s.p[0] / 2.0f; d.hy = s.p[1] / 2.0f; break;
            case CB2_CYLINDER:
            case CB2_CAPSULE: d.hx = s.p[0]; d.hy = s.p[1]; break;
            case CB2_PARTICLE: break;
            case CB2_POLYHEDRON:
                if ((uint64_t)s.vtx_off + s.vtx_cnt > n_verts) return -1;
                d.voff = s.vtx_off;
                d.vcnt = s.vtx_cnt;
                d.hx = 1.0f;      // R is only read by the support-cell lookup
                d.cell_cfg = 0u;  // N = 0: full scan
                break;
            default: return -2;
        }
        T.shapes[i] = d;
    }
    T.verts.resize(n_verts ? n_verts : 1);
    for (uint64_t i = 0; i < n_verts; ++i)
        T.verts[i] = make_float4(verts_xyz[3 * i], verts_xyz[3 * i + 1], verts_xyz[3 * i + 2], 0.0f);
    return 0;
}

template <int F, int V, int E, int RF, int RV>
void run_tier(epa_args C, int tier, int warps, bool pf) {
    C.tier = (uint32_t)tier;
    std::vector<lane_store<F, V, E>> S(warps);
    std::vector<float4> pa((size_t)warps * V * 32);
    // lanes are independent: each runs its state machine until the tier's queue is empty
    for (int w = 0; w < warps; ++w)
        for (unsigned lane = 0; lane < 32; ++lane)
            if (pf) epa_lane_run<F, V, E, RF, RV, true>(C, S[w], pa.data() + (size_t)w * V * 32, lane);
            else epa_lane_run<F, V, E, RF, RV, false>(C, S[w], pa.data() + (size_t)w * V * 32, lane);
}

// the warp-team kernel: every part runs for all warps and lanes of a block before the next part
// starts (what the block barriers of k_epa_team enforce)
template <int W, int F, int V, int E, int RF, int RV>
void run_team_tier(epa_args C, int tier, int blocks) {
    using T = epa_team<W, F, V, E, RF, RV>;
    C.tier = (uint32_t)tier;
    std::vector<typename T::store> S(blocks);
    std::vector<float4> pa((size_t)blocks * V * 32);
    std::vector<typename T::regs> R((size_t)blocks * W * 32);
    for (auto& r : R) T::init(r);
    std::vector<char> done(blocks, 0);
    for (int left = blocks; left > 0;) {
        for (int b = 0; b < blocks; ++b) {
            if (done[b]) continue;
            float4* p = pa.data() + (size_t)b * V * 32;
            typename T::regs* r = R.data() + (size_t)b * W * 32;
#define PART(fn)                                                      \
    for (int w = 0; w < W; ++w)                                       \
        for (unsigned lane = 0; lane < 32; ++lane) T::fn(C, S[b], p, r[w * 32 + lane], w, lane);
            PART(part0) PART(part1) PART(part2) PART(part3) PART(part4) PART(part5)
#undef PART
            bool all = true;
            for (int i = 0; i < W * 32; ++i) all = all && r[i].state == L_EXIT;
            if (all) {
                done[b] = 1;
                --left;
            }
        }
    }
}

}  // namespace

extern "C" {

void sim_force_literal_horizon(int on) { cb2::g_sim_force_literal_horizon = on != 0; }

// GJK3::intersection(FullResolution) over a batch: device GJK source, then the lane tiers.
// caps: 0 -> tier 0 (40,24,16) + tier 1 (64,32,24); 1 -> (32,20,16) + (64,32,24);
// team_w: 0 -> the lane kernel (cb2_epa_lane.cuh); 2..4 -> the warp-team kernel (cb2_epa_team.cuh)
//       2 -> (12,8,8) + (24,14,8) (tiny caps: exercises overflow, dump and resume on small inputs)
// Pairs that outgrow tier 1 come back as status 7 (the CUDA build hands those to the group tiers).
// tier_counts[0..2] = pairs queued for tier 0 / 1 / beyond.
int sim_gjk3_intersection_full(const cb2_shape* shapes, uint32_t n_shapes, const float* verts_xyz,
                               uint64_t n_verts, const uint32_t* l_shape, const uint32_t* r_shape,
                               const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                               const cb2_settings* settings, int caps, int warps, uint32_t dump_cap0, int prefetch, int team_w,
                               cb2_contact* out, uint8_t* status, uint32_t* tier_counts) {
    sim_tables T;
    const int rc = build_tables(shapes, n_shapes, verts_xyz, n_verts, T);
    if (rc) return rc;
    narrow_args A;
    std::memset(&A, 0, sizeof(A));
    A.T.shapes = T.shapes.data();
    A.T.verts = T.verts.data();
    A.l_shape = l_shape;
    A.r_shape = r_shape;
    A.l_t = reinterpret_cast<const float4*>(l_t);
    A.r_t = reinterpret_cast<const float4*>(r_t);
    A.n = n;
    A.n_shapes = n_shapes;
    A.settings = *settings;
    A.out_contact = out;
    A.out_status = status;

    std::vector<uint32_t> lists[EPA_TIERS];
    for (auto& l : lists) l.assign(n ? n : 1, 0u);
    std::vector<float4> simplex((n ? n : 1) * 8);
    epa_counters EC;
    std::memset(&EC, 0, sizeof(EC));
    narrow_ws W;
    std::memset(&W, 0, sizeof(W));
    for (int k = 0; k < EPA_TIERS; ++k) W.list[k] = lists[k].data();
    W.simplex = simplex.data();
    W.EC = &EC;

    // stage 1: GJK::intersect per pair (what k_gjk_hits does)
    for (uint64_t i = 0; i < n; ++i) {
        pair_in p;
        uint8_t st = ST_NONE;
        bool hit = false;
        if (const uint8_t bad = load_pair(A, i, A.l_t, A.r_t, p)) {
            st = bad;
        } else {
            cb2::simplex<true> s;
            s.n = 0;
            const bool inside = gjk_intersect(p, A.settings.max_iterations, s);
            if (inside) {
                hit = true;
                for (int k = 0; k < 4; ++k) {
                    simplex[i * 8 + k] = make_float4(s.v[k].x, s.v[k].y, s.v[k].z, 0.f);
                    simplex[i * 8 + 4 + k] = make_float4(s.a[k].x, s.a[k].y, s.a[k].z, 0.f);
                }
                lists[0][EC.n_list[0]++] = (uint32_t)i;
            }
        }
        if (!hit) {
            store_contact(out, i, CB2_FULL_RESOLUTION, zero_contact());
            status[i] = st;
        }
    }
    tier_counts[0] = EC.n_list[0];

    epa_args C;
    C.A = A;
    C.W = W;
    C.tier = 0;
    std::vector<float4> dump0, dump1;
    auto tiers = [&](auto f0, auto v0, auto e0, auto f1, auto v1, auto e1) {
        constexpr int F0 = decltype(f0)::value, V0 = decltype(v0)::value, E0 = decltype(e0)::value;
        constexpr int F1 = decltype(f1)::value, V1 = decltype(v1)::value, E1 = decltype(e1)::value;
        dump0.resize((size_t)(dump_cap0 ? dump_cap0 : 1) * dump_layout<F0, V0>::SIZE);
        dump1.resize((size_t)(n ? n : 1) * dump_layout<F1, V1>::SIZE);
        C.W.dump[0] = dump0.data();
        C.W.dump_cap[0] = dump_cap0;
        C.W.dump[1] = dump1.data();
        C.W.dump_cap[1] = (uint32_t)n;
        if (team_w == 4) run_team_tier<4, F0, V0, E0, 0, 0>(C, 0, warps);
        else if (team_w == 2) run_team_tier<2, F0, V0, E0, 0, 0>(C, 0, warps);
        else if (team_w == 3) run_team_tier<3, F0, V0, E0, 0, 0>(C, 0, warps);
        else run_tier<F0, V0, E0, 0, 0>(C, 0, warps, prefetch != 0);
        tier_counts[1] = EC.n_list[1];
        if (team_w == 4) run_team_tier<4, F1, V1, E1, F0, V0>(C, 1, warps);
        else if (team_w == 2) run_team_tier<2, F1, V1, E1, F0, V0>(C, 1, warps);
        else if (team_w == 3) run_team_tier<3, F1, V1, E1, F0, V0>(C, 1, warps);
        else run_tier<F1, V1, E1, F0, V0>(C, 1, warps, prefetch != 0);
        tier_counts[2] = EC.n_list[2];
    };
    using std::integral_constant;
#define IC(x) integral_constant<int, x>()
    if (caps == 0) tiers(IC(40), IC(24), IC(16), IC(64), IC(32), IC(24));
    else if (caps == 1) tiers(IC(32), IC(20), IC(16), IC(64), IC(32), IC(24));
    else tiers(IC(12), IC(8), IC(8), IC(24), IC(14), IC(8));
#undef IC
    // what tier 1 could not hold
    for (uint32_t k = 0; k < EC.n_list[2]; ++k) {
        const uint32_t pair = lists[2][k] & ~OVER_RESUME;
        store_contact(out, pair, CB2_FULL_RESOLUTION, zero_contact());
        status[pair] = ST_SCRATCH_OVERFLOW;
    }
    return 0;
}

}  // extern "C"
