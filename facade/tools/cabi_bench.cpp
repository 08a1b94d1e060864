// cabi_bench.cpp -- measurement harness for the C ABI itself (bench.py's end-to-end leg): the loop a C or C++ caller of
// include/cc_b200.h writes to keep the device busy -- submit(k + 2); wait(k) -- over page-locked HOST buffers in the
// reference's 48-byte RawPoint layout, with the results of every push (events, clusters, member lists, packed labels)
// read on the host. No interpreter between the calls: what is timed is the library, the driver and the link.
// Product-side tool: links only libcc_b200.so.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/cc_b200.h"

#define CB_API extern "C" __attribute__((visibility("default")))

// W synchronous warm-up pushes, then K timed pushes of `batch` firings. pts / poses: (warm + steps) * batch firings.
// out[0] seconds of the K timed pushes, out[1] device->host bytes read per push (mean), out[2] checksum of what was read,
// out[3] pushes that took the exact path; marks_ms[steps]: time of every cc_wait return since the clock started.
CB_API int cb_e2e(const cc_config_t* cfg, int rows, const double* robot_from_sensor, int device, int batch, int warm, int steps,
                  const cc_raw_point_t* pts, const double* poses, int label_prefetch, double* out, double* marks_ms, float* slot_ms,
                  char* err)
{
    cc_handle_t* h = nullptr;
    auto fail = [&](const char* what) -> int
    {
        if (err)
            std::snprintf(err, 256, "%s: %s", what, h ? cc_last_error(h) : "no handle");
        if (h)
            cc_destroy(h);
        return 1;
    };
    if (cc_create(device, batch, &h) != CC_OK)
        return fail("cc_create");
    if (cc_set_config(h, cfg) != CC_OK || cc_reset(h, rows) != CC_OK || cc_set_robot_from_sensor(h, robot_from_sensor) != CC_OK ||
        cc_set_label_prefetch(h, label_prefetch) != CC_OK)
        return fail("configure");
    const size_t stride = static_cast<size_t>(batch) * rows;
    for (int s = 0; s < warm; s++)
        if (cc_push_firings(h, batch, rows, pts + s * stride, poses + static_cast<size_t>(s) * batch * 12) != CC_OK)
            return fail("cc_push_firings");
    uint64_t d2h = 0, checksum = 0, exact = 0;
    auto consume = [&]() -> bool
    {
        cc_batch_info_t info;
        const cc_column_event_t* ev = nullptr;
        const cc_cluster_t* cl = nullptr;
        const cc_cluster_point_t* cp = nullptr;
        if (cc_get_batch_info(h, &info) != CC_OK || cc_get_result_views(h, &ev, &cl, &cp) != CC_OK)
            return false;
        for (int i = 0; i < info.n_events; i++)
            checksum += static_cast<uint64_t>(ev[i].to_gcol);
        for (int i = 0; i < info.n_clusters; i++)
            checksum += cl[i].stamp & 0xffff;
        if (info.n_cluster_points > 0)
            checksum += static_cast<uint64_t>(cp[info.n_cluster_points - 1].gcol);
        d2h += static_cast<uint64_t>(info.n_clusters) * sizeof(cc_cluster_t) + static_cast<uint64_t>(info.n_cluster_points) * sizeof(cc_cluster_point_t) + 512;
        if (label_prefetch)
        {
            const uint8_t* labels = nullptr;
            int ncols = 0;
            if (cc_get_column_labels(h, &labels, &ncols) != CC_OK)
                return false;
            if (ncols > 0)
                checksum += labels[static_cast<size_t>(ncols) * rows * 4 - 4];
            d2h += static_cast<uint64_t>(ncols) * rows * 4 + static_cast<uint64_t>(ncols) * 8;
        }
        exact += info.used_exact_path ? 1 : 0;
        return true;
    };
    if (slot_ms)
        cc_debug_slot_base(h);
    const auto t0 = std::chrono::steady_clock::now();
    // two pushes in flight and a third one staged: the host->device copy of push k + 2 overlaps the kernels of k and k + 1
    for (int s = warm; s < warm + steps && s < warm + 2; s++)
        if (cc_submit_firings(h, batch, rows, pts + s * stride, poses + static_cast<size_t>(s) * batch * 12) != CC_OK)
            return fail("cc_submit_firings");
    for (int s = warm; s < warm + steps; s++)
    {
        if (s + 2 < warm + steps &&
            cc_submit_firings(h, batch, rows, pts + (s + 2) * stride, poses + static_cast<size_t>(s + 2) * batch * 12) != CC_OK)
            return fail("cc_submit_firings");
        if (cc_wait(h) != CC_OK)
            return fail("cc_wait");
        if (marks_ms)
            marks_ms[s - warm] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (slot_ms)
            cc_debug_slot_times(h, s % 2, slot_ms + 5 * (s - warm)); // the slots alternate from the first push on
        if (!consume())
            return fail("results");
    }
    out[0] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    out[1] = steps > 0 ? static_cast<double>(d2h) / steps : 0.0;
    out[2] = static_cast<double>(checksum % 1000000007ull);
    out[3] = static_cast<double>(exact);
    cc_destroy(h);
    return 0;
}

// Latency mode: `n` synchronous pushes of `batch` firings (cc_push_firings: host buffers in, results on the host when the
// call returns, read here), one after the other. call_us[n]: wall time of every push incl. reading its results;
// device_us[n]: the device time of the push as the library reports it (cc_batch_info_t::device_ms).
CB_API int cb_latency(const cc_config_t* cfg, int rows, const double* robot_from_sensor, int device, int batch, int n,
                      const cc_raw_point_t* pts, const double* poses, int label_prefetch, double* call_us, double* device_us,
                      double* out, char* err)
{
    cc_handle_t* h = nullptr;
    auto fail = [&](const char* what) -> int
    {
        if (err)
            std::snprintf(err, 256, "%s: %s", what, h ? cc_last_error(h) : "no handle");
        if (h)
            cc_destroy(h);
        return 1;
    };
    if (cc_create(device, batch, &h) != CC_OK)
        return fail("cc_create");
    if (cc_set_config(h, cfg) != CC_OK || cc_reset(h, rows) != CC_OK || cc_set_robot_from_sensor(h, robot_from_sensor) != CC_OK ||
        cc_set_label_prefetch(h, label_prefetch) != CC_OK)
        return fail("configure");
    const size_t stride = static_cast<size_t>(batch) * rows;
    uint64_t checksum = 0, launches = 0;
    for (int s = 0; s < n; s++)
    {
        const auto t0 = std::chrono::steady_clock::now();
        if (cc_push_firings(h, batch, rows, pts + s * stride, poses + static_cast<size_t>(s) * batch * 12) != CC_OK)
            return fail("cc_push_firings");
        cc_batch_info_t info;
        const cc_column_event_t* ev = nullptr;
        const cc_cluster_t* cl = nullptr;
        const cc_cluster_point_t* cp = nullptr;
        if (cc_get_batch_info(h, &info) != CC_OK || cc_get_result_views(h, &ev, &cl, &cp) != CC_OK)
            return fail("results");
        for (int i = 0; i < info.n_events; i++)
            checksum += static_cast<uint64_t>(ev[i].to_gcol);
        for (int i = 0; i < info.n_clusters; i++)
            checksum += cl[i].stamp & 0xffff;
        if (info.n_cluster_points > 0)
            checksum += static_cast<uint64_t>(cp[info.n_cluster_points - 1].gcol);
        if (label_prefetch)
        {
            const uint8_t* labels = nullptr;
            int ncols = 0;
            if (cc_get_column_labels(h, &labels, &ncols) != CC_OK)
                return fail("labels");
            if (ncols > 0)
                checksum += labels[static_cast<size_t>(ncols) * rows * 4 - 4];
        }
        call_us[s] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        device_us[s] = 1e3 * info.device_ms;
        launches = static_cast<uint64_t>(info.gpu_launches);
    }
    out[0] = static_cast<double>(checksum % 1000000007ull);
    out[1] = static_cast<double>(launches);
    cc_destroy(h);
    return 0;
}
