// cb2_multi.cu — one handle over several GPUs of one box (the cb2_create(devices, n) of SURVEY §8(b)).
//
// The narrow phase shards by pairs (independent work, no exchange: DESIGN §6), the broad phase by
// the sorted position of a pair's later box (cb2_sap3_find_collider_pairs_shard_dev: the shards
// concatenate to the reference's Vec byte for byte).  A cb2_multi owns one cb2_ctx per device and
// one host thread per device per call; every thread drives its device's own host-buffer pipeline
// on a contiguous slice of the caller's arrays, so a single cb2_multi_gjk3_intersection call from
// Rust keeps all GPUs busy.  The shape table is uploaded to every device (23 MB for the headline
// table: a one-off).  Worker threads pin themselves to the CPUs of their GPU's NUMA node (sysfs),
// so that staging copies do not cross the socket interconnect.
//
// Built only on the public single-device C ABI: nothing here touches a kernel.
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/collision_b200.h"

struct cb2_multi {
    std::vector<cb2_ctx*> ctx;
    std::vector<int> device;
    std::vector<cpu_set_t> cpus;   // CPUs local to each device (empty set: leave the thread alone)
    std::vector<char> has_cpus;
    std::string err;
    // broad-phase scratch per device
    std::vector<void*> d_boxes, d_perm, d_pairs;
    std::vector<uint64_t> boxes_cap, pairs_cap;
};

namespace {

// CPUs of the NUMA node the device hangs off: /sys/bus/pci/devices/<bdf>/local_cpulist
bool local_cpus(int device, cpu_set_t* set) {
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof(bdf), device) != cudaSuccess) return false;
    for (char* p = bdf; *p; ++p) *p = (char)tolower(*p);
    const std::string path = std::string("/sys/bus/pci/devices/") + bdf + "/local_cpulist";
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return false;
    char line[4096] = {0};
    const bool ok = std::fgets(line, sizeof(line), f) != nullptr;
    std::fclose(f);
    if (!ok) return false;
    CPU_ZERO(set);
    int n = 0;
    for (char* tok = std::strtok(line, ",\n"); tok; tok = std::strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        if (std::sscanf(tok, "%d-%d", &a, &b) == 2) {
            for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, set); ++n; }
        } else if (std::sscanf(tok, "%d", &a) == 1 && a < CPU_SETSIZE) {
            CPU_SET(a, set);
            ++n;
        }
    }
    return n > 0;
}

// slice i of n items over k devices: contiguous, sizes differ by at most one
inline void slice(uint64_t n, int k, int i, uint64_t* lo, uint64_t* hi) {
    const uint64_t base = n / (uint64_t)k, rem = n % (uint64_t)k;
    *lo = (uint64_t)i * base + std::min<uint64_t>((uint64_t)i, rem);
    *hi = *lo + base + ((uint64_t)i < rem ? 1 : 0);
}

template <class F>
int on_every_device(cb2_multi* m, F f) {
    const int k = (int)m->ctx.size();
    std::vector<int> rc(k, CB2_OK);
    std::vector<std::thread> th;
    th.reserve(k);
    for (int i = 0; i < k; ++i)
        th.emplace_back([&, i] {
            if (m->has_cpus[i]) sched_setaffinity(0, sizeof(cpu_set_t), &m->cpus[i]);
            cudaSetDevice(m->device[i]);
            rc[i] = f(i);
        });
    for (auto& t : th) t.join();
    for (int i = 0; i < k; ++i)
        if (rc[i] != CB2_OK) {
            const char* e = cb2_last_error(m->ctx[i]);
            m->err = "device " + std::to_string(m->device[i]) + ": " + (e ? e : "");
            return rc[i];
        }
    return CB2_OK;
}

}  // namespace

extern "C" {

int cb2_create_multi(const int* devices, int n_devices, cb2_multi** out) {
    if (!out) return CB2_ERR_ARG;
    *out = nullptr;
    if (!devices || n_devices < 1) return CB2_ERR_ARG;
    cb2_multi* m = new cb2_multi();
    for (int i = 0; i < n_devices; ++i) {
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) {
                cb2_destroy_multi(m);
                return CB2_ERR_ARG;  // one context per device
            }
        cb2_ctx* c = nullptr;
        const int rc = cb2_create(devices[i], &c);
        if (rc != CB2_OK) {
            cb2_destroy_multi(m);
            return rc;  // cb2_last_error(NULL) has the reason
        }
        m->ctx.push_back(c);
        m->device.push_back(devices[i]);
        cpu_set_t s;
        const bool ok = local_cpus(devices[i], &s);
        if (!ok) CPU_ZERO(&s);
        m->cpus.push_back(s);
        m->has_cpus.push_back(ok ? 1 : 0);
    }
    m->d_boxes.assign(n_devices, nullptr);
    m->d_perm.assign(n_devices, nullptr);
    m->d_pairs.assign(n_devices, nullptr);
    m->boxes_cap.assign(n_devices, 0);
    m->pairs_cap.assign(n_devices, 0);
    *out = m;
    return CB2_OK;
}

void cb2_destroy_multi(cb2_multi* m) {
    if (!m) return;
    for (size_t i = 0; i < m->ctx.size(); ++i) {
        cudaSetDevice(m->device[i]);
        if (i < m->d_boxes.size()) {
            if (m->d_boxes[i]) cudaFree(m->d_boxes[i]);
            if (m->d_perm[i]) cudaFree(m->d_perm[i]);
            if (m->d_pairs[i]) cudaFree(m->d_pairs[i]);
        }
        cb2_destroy(m->ctx[i]);
    }
    delete m;
}

int cb2_multi_device_count(const cb2_multi* m) { return m ? (int)m->ctx.size() : 0; }
cb2_ctx* cb2_multi_ctx(cb2_multi* m, int i) {
    return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr;
}
const char* cb2_multi_last_error(const cb2_multi* m) { return m ? m->err.c_str() : cb2_last_error(nullptr); }

int cb2_multi_shapes_upload(cb2_multi* m, const cb2_shape* shapes, uint32_t n_shapes,
                            const float* verts_xyz, uint64_t n_verts) {
    if (!m) return CB2_ERR_ARG;
    return on_every_device(m, [&](int i) { return cb2_shapes_upload(m->ctx[i], shapes, n_shapes, verts_xyz, n_verts); });
}

int cb2_multi_gjk3_intersection(cb2_multi* m, uint32_t strategy, const cb2_settings* settings,
                                const uint32_t* l_shape, const uint32_t* r_shape, const cb2_xform* l_t,
                                const cb2_xform* r_t, uint64_t n, cb2_contact* out, uint8_t* status) {
    if (!m) return CB2_ERR_ARG;
    const int k = (int)m->ctx.size();
    return on_every_device(m, [&](int i) {
        uint64_t lo, hi;
        slice(n, k, i, &lo, &hi);
        return cb2_gjk3_intersection(m->ctx[i], strategy, settings, l_shape + lo, r_shape + lo, l_t + lo,
                                     r_t + lo, hi - lo, out + lo, status + lo);
    });
}

int cb2_multi_gjk3_distance(cb2_multi* m, const cb2_settings* settings, const uint32_t* l_shape,
                            const uint32_t* r_shape, const cb2_xform* l_t, const cb2_xform* r_t, uint64_t n,
                            float* out, uint8_t* status) {
    if (!m) return CB2_ERR_ARG;
    const int k = (int)m->ctx.size();
    return on_every_device(m, [&](int i) {
        uint64_t lo, hi;
        slice(n, k, i, &lo, &hi);
        return cb2_gjk3_distance(m->ctx[i], settings, l_shape + lo, r_shape + lo, l_t + lo, r_t + lo, hi - lo,
                                 out + lo, status + lo);
    });
}

int cb2_multi_gjk3_time_of_impact(cb2_multi* m, const cb2_settings* settings, const uint32_t* l_shape,
                                  const uint32_t* r_shape, const cb2_xform* l_t0, const cb2_xform* l_t1,
                                  const cb2_xform* r_t0, const cb2_xform* r_t1, uint64_t n, uint32_t iter_cap,
                                  cb2_contact* out, uint8_t* status) {
    if (!m) return CB2_ERR_ARG;
    const int k = (int)m->ctx.size();
    return on_every_device(m, [&](int i) {
        uint64_t lo, hi;
        slice(n, k, i, &lo, &hi);
        return cb2_gjk3_time_of_impact(m->ctx[i], settings, l_shape + lo, r_shape + lo, l_t0 + lo, l_t1 + lo,
                                       r_t0 + lo, r_t1 + lo, hi - lo, iter_cap, out + lo, status + lo);
    });
}

// SweepAndPrune3::find_collider_pairs over all devices: every device sorts the whole table (cheaper
// than exchanging it) and sweeps, but reports only its shard of the pair list; the shards are
// concatenated in order, which is the reference's Vec.
int cb2_multi_sap3_find_collider_pairs(cb2_multi* m, const cb2_aabb3* aabbs, uint64_t n, uint32_t* axis_inout,
                                       uint32_t* perm_out, uint32_t* pairs_out, uint64_t cap,
                                       uint64_t* n_pairs) {
    if (!m || !axis_inout || !n_pairs) return CB2_ERR_ARG;
    *n_pairs = 0;
    if (*axis_inout > 2) {
        m->err = "sweep axis must be 0, 1 or 2";
        return CB2_ERR_ARG;
    }
    if (n == 0) return CB2_OK;
    if (!aabbs || !perm_out || (!pairs_out && cap)) {
        m->err = "null argument";
        return CB2_ERR_ARG;
    }
    if (n == 1) {  // sweep_prune.rs:83-85: early return, axis untouched
        perm_out[0] = 0;
        return CB2_OK;
    }
    const int k = (int)m->ctx.size();
    const uint32_t axis = *axis_inout;
    std::vector<uint64_t> count(k, 0);
    std::vector<std::vector<uint32_t>> shard(k);
    int rc = on_every_device(m, [&](int i) -> int {
        auto fail = [&](const char*) { return CB2_ERR_CUDA; };
        if (m->boxes_cap[i] < n) {
            if (m->d_boxes[i]) cudaFree(m->d_boxes[i]);
            if (m->d_perm[i]) cudaFree(m->d_perm[i]);
            m->d_boxes[i] = m->d_perm[i] = nullptr;
            m->boxes_cap[i] = 0;
            if (cudaMalloc(&m->d_boxes[i], n * sizeof(cb2_aabb3)) != cudaSuccess) return fail("cudaMalloc");
            if (cudaMalloc(&m->d_perm[i], n * 4) != cudaSuccess) return fail("cudaMalloc");
            m->boxes_cap[i] = n;
        }
        if (cudaMemcpy(m->d_boxes[i], aabbs, n * sizeof(cb2_aabb3), cudaMemcpyHostToDevice) != cudaSuccess)
            return fail("cudaMemcpy");
        uint64_t want = std::max<uint64_t>(m->pairs_cap[i], 16 * n / (uint64_t)k + 1024);
        for (int attempt = 0; attempt < 2; ++attempt) {
            if (m->pairs_cap[i] < want) {
                if (m->d_pairs[i]) cudaFree(m->d_pairs[i]);
                m->d_pairs[i] = nullptr;
                m->pairs_cap[i] = 0;
                if (cudaMalloc(&m->d_pairs[i], want * 8) != cudaSuccess) return fail("cudaMalloc");
                m->pairs_cap[i] = want;
            }
            const int r = cb2_sap3_find_collider_pairs_shard_dev(
                m->ctx[i], static_cast<cb2_aabb3*>(m->d_boxes[i]), n, axis, (uint32_t)i, (uint32_t)k,
                static_cast<uint32_t*>(m->d_perm[i]), static_cast<uint32_t*>(m->d_pairs[i]), m->pairs_cap[i],
                &count[i], nullptr);
            if (r == CB2_ERR_NOSPC && attempt == 0) {
                want = count[i];
                continue;
            }
            if (r != CB2_OK) return r;
            break;
        }
        if (cudaDeviceSynchronize() != cudaSuccess) return fail("sync");
        shard[i].resize(2 * count[i]);
        if (count[i] && cudaMemcpy(shard[i].data(), m->d_pairs[i], count[i] * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail("cudaMemcpy");
        if (i == 0 && cudaMemcpy(perm_out, m->d_perm[i], n * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail("cudaMemcpy");
        return CB2_OK;
    });
    if (rc != CB2_OK) return rc;
    uint64_t total = 0;
    for (int i = 0; i < k; ++i) total += count[i];
    *n_pairs = total;
    if (total > cap) {
        m->err = "pairs_out too small; *n_pairs holds the required capacity";
        return CB2_ERR_NOSPC;  // the axis is left alone so the caller can retry
    }
    uint64_t off = 0;
    for (int i = 0; i < k; ++i) {
        if (count[i]) std::memcpy(pairs_out + 2 * off, shard[i].data(), count[i] * 8);
        off += count[i];
    }
    *axis_inout = 0;  // compute_axis sees zero sums (sweep_prune.rs:130-133, 254-277)
    return CB2_OK;
}

}  // extern "C"
