// rsp_cluster.cpp -- S10 + S11 on the host: the two clustering stages of fun_process_single_frame.m:302-407.
//
// The reference grows each cluster with a FIFO breadth-first search seeded in index order (fsf:313-326,
// :366-379).  The link relation is symmetric, so what that search computes is exactly the connected
// components of the link graph, numbered by their smallest member index; the power-weighted sums are then
// taken in index order (fsf:333-347), not in visiting order.  This file computes the same labelling with a
// union-find over a range-sorted sweep (only detections within max_range_sep can be linked), which is
// O(n * window) instead of the O(n^2) of the literal search -- the literal search was the largest host cost
// of the frame path (89 us for 205 detections vs a 58 us device chain).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

#include "rsp.h"

namespace {

// ids[i] in 1..n_clusters = connected component of i, numbered by smallest member (== the reference's BFS ids).
// Points are (key, u, w) with the link relation |dkey| <= sep_key && |du| <= sep_u && |dw| <= sep_w (sep_w < 0: w unused).
// Dense scenes (64 targets: 1600 detections in clumps of 25) made this the largest host cost of the frame path, so the
// sweep avoids both a comparison sort and unpredictable branches: the points are counting-sorted into cells a quarter of
// sep_key wide (two linked points are at most four cells apart), every point meets the rest of its cell and the next four
// cells -- one contiguous stretch of the compact arrays --, the three gates are evaluated without branches and only the
// linked candidates reach the union-find.  The gates themselves are the reference's (fsf:319-321, :372-373).
struct SweepPoint { double key, u, w; int idx; };

template <typename Get> std::vector<int> component_ids(int n, Get get, double sep_key, double sep_u, double sep_w, int* n_clusters) {
    constexpr int kReach = 4;
    std::vector<SweepPoint> in(n);
    double kmin = 0, kmax = 0;
    bool any = false;
    for (int i = 0; i < n; ++i) {
        in[i] = get(i);
        const double k = in[i].key;
        if (std::isfinite(k)) { kmin = any ? std::min(kmin, k) : k; kmax = any ? std::max(kmax, k) : k; any = true; }
    }
    // cell width: sep_key / 4 with a margin far above the rounding of the quotient, widened until there are at most 4 n cells
    double width = (sep_key > 0 ? sep_key : 0.0) * (1.0 + 1e-9) / kReach;
    width = std::max(width, (kmax - kmin) / (4.0 * n));
    if (!(width > 0) || !std::isfinite(width)) width = 1.0;          // all keys equal (or none finite): one cell
    const int n_cells = (int)std::min((kmax - kmin) / width, 4.0 * n + 8.0) + 1;
    std::vector<int> cell(n), start(n_cells + kReach + 2, 0);
    for (int i = 0; i < n; ++i) {
        const double k = in[i].key;                                   // a non-finite key links to nothing; any cell will do
        cell[i] = std::isfinite(k) ? std::min((int)((k - kmin) / width), n_cells - 1) : 0;
        ++start[cell[i] + 1];
    }
    for (int c = 0; c < n_cells; ++c) start[c + 1] += start[c];
    for (int c = n_cells + 1; c < (int)start.size(); ++c) start[c] = n;
    std::vector<double> key(n), u(n), w(n);
    std::vector<int> idx(n), fill(start.begin(), start.begin() + n_cells);
    for (int i = 0; i < n; ++i) {                                     // stable: index order inside a cell
        const int pos = fill[cell[i]]++;
        key[pos] = in[i].key; u[pos] = in[i].u; w[pos] = in[i].w; idx[pos] = i;
    }
    std::vector<int> parent(n), cand(n);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int x) {
        while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
        return x;
    };
    const bool use_w = sep_w >= 0;
    for (int c = 0; c < n_cells; ++c) {
        const int end = start[c + kReach + 1];
        for (int a = start[c]; a < start[c + 1]; ++a) {
            const double ka = key[a], ua = u[a], wa = w[a];
            int m = 0;
            for (int b = a + 1; b < end; ++b) {
                const bool ok = (std::fabs(ka - key[b]) <= sep_key) & (std::fabs(ua - u[b]) <= sep_u) &
                                (!use_w | (std::fabs(wa - w[b]) <= sep_w));
                cand[m] = idx[b];
                m += ok;
            }
            if (m == 0) continue;
            int ra = find(idx[a]);
            for (int j = 0; j < m; ++j) {
                const int rb = find(cand[j]);
                if (ra != rb) { const int lo = std::min(ra, rb); parent[std::max(ra, rb)] = lo; ra = lo; }
            }
        }
    }
    std::vector<int> ids(n), label(n, 0);
    int cur = 0;
    for (int i = 0; i < n; ++i) {                // index order => a component is numbered when its smallest member is met
        const int r = find(i);
        if (label[r] == 0) label[r] = ++cur;     // fsf:315 / :368
        ids[i] = label[r];
    }
    *n_clusters = cur;
    return ids;
}

}  // namespace

extern "C" int rsp_cluster(const rsp_detection* dets, int32_t n, const rsp_cluster_params* cp, rsp_target* stage1,
                           int32_t* n_stage1, rsp_target* final_targets, int32_t* n_final) {
    if (n < 0 || !cp || !n_final || (n > 0 && (!dets || !final_targets))) return RSP_ERR_INVALID_ARG;
    if (n_stage1) *n_stage1 = 0;
    *n_final = 0;
    if (n == 0) return RSP_OK;                   // fsf:305-308, :358-361 -> []
    // ---- stage 1: R/V/Angle gates, power-weighted mean, Power = sum (fsf:313-351)
    int nc1 = 0;
    std::vector<int> ids1 = component_ids(n, [&](int i) { return SweepPoint{dets[i].range, dets[i].velocity, dets[i].angle, 0}; },
                                          cp->max_range_sep, cp->max_vel_sep, cp->max_angle_sep, &nc1);
    std::vector<rsp_target> t1(nc1, rsp_target{0, 0, 0, 0});
    for (int i = 0; i < n; ++i) {                // one pass; every cluster still sums its members in index order
        rsp_target& t = t1[ids1[i] - 1];
        const double pw = (double)dets[i].power;
        t.power += pw;
        t.range += dets[i].range * pw;
        t.velocity += dets[i].velocity * pw;
        t.angle += dets[i].angle * pw;
    }
    for (auto& t : t1) { t.range /= t.power; t.velocity /= t.power; t.angle /= t.power; }
    if (n_stage1) *n_stage1 = nc1;
    if (stage1) std::memcpy(stage1, t1.data(), (size_t)nc1 * sizeof(rsp_target));
    // ---- stage 2: R/V gates, winner takes all (fsf:366-406)
    int nc2 = 0;
    std::vector<int> ids2 = component_ids(nc1, [&](int i) { return SweepPoint{t1[i].range, t1[i].velocity, 0.0, 0}; },
                                          cp->max_range_sep, cp->max_vel_sep, -1.0, &nc2);
    std::vector<int> win(nc2, -1);
    for (int i = 0; i < nc1; ++i) {
        int& w = win[ids2[i] - 1];
        if (w < 0 || t1[i].power > t1[w].power) w = i;                                  // first max, fsf:399
    }
    for (int cidx = 0; cidx < nc2; ++cidx) final_targets[cidx] = t1[win[cidx]];
    *n_final = nc2;
    return RSP_OK;
}
