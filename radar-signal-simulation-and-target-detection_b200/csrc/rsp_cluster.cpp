// rsp_cluster.cpp -- S10 + S11 on the host: the two order-dependent BFS clustering stages of
// fun_process_single_frame.m:302-407.  n is a few hundred detections, so this is O(n^2) host work;
// cluster ids follow the reference's visiting order (FIFO queue seeded in index order), which is
// what makes final_targets(1) etc. reproducible.
#include <cmath>
#include <cstring>
#include <deque>
#include <vector>

#include "rsp.h"

namespace {

template <typename Linked> std::vector<int> bfs_ids(int n, Linked linked, int* n_clusters) {
    std::vector<int> ids(n, 0);
    int cur = 0;
    std::deque<int> queue;
    for (int i = 0; i < n; ++i) {
        if (ids[i] != 0) continue;
        ++cur;                                   // fsf:315 / :368
        queue.clear();
        queue.push_back(i);
        while (!queue.empty()) {
            const int k = queue.front();         // fsf:318-319 (FIFO, duplicates allowed)
            queue.pop_front();
            if (ids[k] != 0) continue;
            ids[k] = cur;
            for (int j = 0; j < n; ++j)
                if (ids[j] == 0 && linked(k, j)) queue.push_back(j);
        }
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
    std::vector<int> ids1 = bfs_ids(n, [&](int k, int j) {
        return std::fabs(dets[k].range - dets[j].range) <= cp->max_range_sep &&
               std::fabs(dets[k].velocity - dets[j].velocity) <= cp->max_vel_sep &&
               std::fabs(dets[k].angle - dets[j].angle) <= cp->max_angle_sep;
    }, &nc1);
    std::vector<rsp_target> t1(nc1);
    for (int cidx = 1; cidx <= nc1; ++cidx) {
        double tot = 0, sr = 0, sv = 0, sa = 0;
        for (int i = 0; i < n; ++i)
            if (ids1[i] == cidx) {
                const double pw = (double)dets[i].power;
                tot += pw;
                sr += dets[i].range * pw;
                sv += dets[i].velocity * pw;
                sa += dets[i].angle * pw;
            }
        t1[cidx - 1] = rsp_target{sr / tot, sv / tot, sa / tot, tot};
    }
    if (n_stage1) *n_stage1 = nc1;
    if (stage1) std::memcpy(stage1, t1.data(), (size_t)nc1 * sizeof(rsp_target));
    // ---- stage 2: R/V gates, winner takes all (fsf:366-406)
    int nc2 = 0;
    std::vector<int> ids2 = bfs_ids(nc1, [&](int k, int j) {
        return std::fabs(t1[k].range - t1[j].range) <= cp->max_range_sep &&
               std::fabs(t1[k].velocity - t1[j].velocity) <= cp->max_vel_sep;
    }, &nc2);
    for (int cidx = 1; cidx <= nc2; ++cidx) {
        int win = -1;
        for (int i = 0; i < nc1; ++i)
            if (ids2[i] == cidx && (win < 0 || t1[i].power > t1[win].power)) win = i;   // first max, fsf:399
        final_targets[cidx - 1] = t1[win];
    }
    *n_final = nc2;
    return RSP_OK;
}
