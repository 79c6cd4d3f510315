// Locality ordering of the cells for batched propagation.
//
// The gather kernel reads, for a block of consecutive destination rows, the iterate rows of all their sources. In the
// caller's cell order those sources are scattered over the whole iterate, which is fine while a column tile of X
// stays in L2 (n = 50k) but sends every gather to HBM once it does not (n >= 250k: measured 7.4e10 -> 1.24e11
// cell-iterations/s with the ordering below, profiles/k1_sweep3_rcm_r01.jsonl). Cuthill-McKee (breadth-first levels of the
// symmetrised graph, neighbours by ascending degree, started from a pseudo-peripheral cell) puts graph neighbours
// close together, so the rows gathered by one wave of CTAs form a narrow window of the permuted iterate.
//
// Built once per plan on the host (O(nnz log k)); rejected when it does not shrink the mean |pos(i) - pos(j)| of the
// stored entries by at least 20% (an input that is already well ordered keeps its order).
#include <algorithm>
#include <numeric>
#include <queue>
#include <stdlib.h>

#include "common.cuh"

namespace cy2p {

static void cuthill_mckee(int32_t n, const std::vector<int32_t> &out_ptr, const std::vector<int32_t> &out_idx,
                          const std::vector<int32_t> &in_ptr, const std::vector<Pair> &in_pairs,
                          std::vector<int32_t> &order) {
    std::vector<int32_t> deg(n);
    for (int32_t v = 0; v < n; ++v) deg[v] = (out_ptr[v + 1] - out_ptr[v]) + (in_ptr[v + 1] - in_ptr[v]);
    std::vector<uint8_t> seen(n, 0);
    std::vector<int32_t> level_nodes, nbrs;
    order.clear();
    order.reserve(n);
    auto bfs = [&](int32_t start, std::vector<int32_t> *out, std::vector<uint8_t> &mark) -> int32_t {
        // returns the last node reached; appends the visit order to *out when given
        std::vector<int32_t> q;
        q.push_back(start);
        mark[start] = 1;
        size_t head = 0;
        while (head < q.size()) {
            const int32_t v = q[head++];
            nbrs.clear();
            for (int32_t p = out_ptr[v]; p < out_ptr[v + 1]; ++p) {
                const int32_t u = out_idx[p];
                if (!mark[u]) {
                    mark[u] = 1;
                    nbrs.push_back(u);
                }
            }
            for (int32_t p = in_ptr[v]; p < in_ptr[v + 1]; ++p) {
                const int32_t u = in_pairs[p].src;
                if (!mark[u]) {
                    mark[u] = 1;
                    nbrs.push_back(u);
                }
            }
            std::sort(nbrs.begin(), nbrs.end(), [&](int32_t a, int32_t b) { return deg[a] != deg[b] ? deg[a] < deg[b] : a < b; });
            q.insert(q.end(), nbrs.begin(), nbrs.end());
        }
        if (out) out->insert(out->end(), q.begin(), q.end());
        return q.back();
    };
    std::vector<uint8_t> tmp(n, 0);
    for (int32_t v = 0; v < n; ++v) {
        if (seen[v]) continue;
        // pseudo-peripheral start: the last node of a BFS from v (restricted to v's component)
        std::vector<int32_t> comp;
        const int32_t far = bfs(v, &comp, tmp);
        for (int32_t u : comp) tmp[u] = 0;
        bfs(far, &order, seen);
    }
}

static double mean_span(int32_t n, const std::vector<int32_t> &in_ptr, const std::vector<Pair> &in_pairs,
                        const int32_t *pos) {
    double s = 0.0;
    for (int32_t j = 0; j < n; ++j)
        for (int32_t p = in_ptr[j]; p < in_ptr[j + 1]; ++p) {
            const int64_t a = pos ? pos[j] : j, b = pos ? pos[in_pairs[p].src] : in_pairs[p].src;
            s += (double)(a > b ? a - b : b - a);
        }
    return in_pairs.empty() ? 0.0 : s / (double)in_pairs.size();
}

// Builds the ordering and the permuted T^T on the plan's device. Synchronises `st`. On any failure the plan is left
// in the rejected state (the identity path keeps working).
int32_t ensure_reordered(cy2p_plan *plan, cudaStream_t st) {
    if (plan->reorder_state != 0) return CY2P_OK;
    plan->reorder_state = -1;
    const char *ev = getenv("CY2P_REORDER");
    if (ev && *ev && atoi(ev) == 0) return CY2P_OK;
    const int32_t n = plan->n;
    const size_t nnz = (size_t)plan->nnz;
    if (nnz == 0) return CY2P_OK;
    std::vector<int32_t> out_ptr((size_t)n + 1), out_idx(nnz), in_ptr((size_t)n + 1);
    std::vector<Pair> in_pairs(nnz);
    CY2P_CUDA(cudaMemcpyAsync(out_ptr.data(), plan->d_indptr, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost, st));
    CY2P_CUDA(cudaMemcpyAsync(out_idx.data(), plan->d_indices, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost, st));
    CY2P_CUDA(cudaMemcpyAsync(in_ptr.data(), plan->d_t_indptr, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost, st));
    CY2P_CUDA(cudaMemcpyAsync(in_pairs.data(), plan->d_t_pairs, sizeof(Pair) * nnz, cudaMemcpyDeviceToHost, st));
    CY2P_CUDA(cudaStreamSynchronize(st));

    std::vector<int32_t> order;
    cuthill_mckee(n, out_ptr, out_idx, in_ptr, in_pairs, order);
    if ((int32_t)order.size() != n) return CY2P_OK;
    std::vector<int32_t> pos(n);
    for (int32_t q = 0; q < n; ++q) pos[order[q]] = q;
    plan->span_identity = mean_span(n, in_ptr, in_pairs, nullptr);
    plan->span_reordered = mean_span(n, in_ptr, in_pairs, pos.data());
    if (!(plan->span_reordered < 0.8 * plan->span_identity)) return CY2P_OK;  // already well ordered

    std::vector<int32_t> r_ptr((size_t)n + 1);
    std::vector<Pair> pp(nnz), op(nnz);
    r_ptr[0] = 0;
    for (int32_t q = 0; q < n; ++q) r_ptr[q + 1] = r_ptr[q] + (in_ptr[order[q] + 1] - in_ptr[order[q]]);
    for (int32_t q = 0; q < n; ++q) {
        const int32_t j = order[q];
        const int32_t len = in_ptr[j + 1] - in_ptr[j];
        Pair *dpp = pp.data() + r_ptr[q], *dop = op.data() + r_ptr[q];
        for (int32_t t = 0; t < len; ++t) {
            const Pair e = in_pairs[in_ptr[j] + t];
            dop[t] = e;  // ascending original source, as in T^T
            dpp[t] = Pair{pos[e.src], e.val};
        }
        std::sort(dpp, dpp + len, [](const Pair &a, const Pair &b) { return a.src < b.src; });  // ascending addresses
    }
    auto upload = [&](void **dst, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes ? bytes : 1);
        if (e != cudaSuccess) return e;
        plan->bytes_owned += bytes;
        return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, st);
    };
    cudaError_t e = upload((void **)&plan->d_order, order.data(), sizeof(int32_t) * (size_t)n);
    if (e == cudaSuccess) e = upload((void **)&plan->d_r_indptr, r_ptr.data(), sizeof(int32_t) * ((size_t)n + 1));
    if (e == cudaSuccess) e = upload((void **)&plan->d_r_pairs_pp, pp.data(), sizeof(Pair) * nnz);
    if (e == cudaSuccess) e = upload((void **)&plan->d_r_pairs_op, op.data(), sizeof(Pair) * nnz);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the host vectors go out of scope
    if (e != cudaSuccess) {
        set_error("reorder: %s", cudaGetErrorString(e));
        return CY2P_ERR_CUDA;
    }
    plan->reorder_state = 1;
    return CY2P_OK;
}

}  // namespace cy2p
