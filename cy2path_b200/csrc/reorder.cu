// Locality ordering of the cells for batched propagation.
//
// The gather kernel reads, for a block of consecutive destination rows, the iterate rows of all their sources. In the
// caller's cell order those sources are scattered over the whole iterate, which is fine while a column tile of X
// stays in L2 (n = 50k) but sends every gather to HBM once it does not (n >= 250k: measured 7.4e10 -> 1.24e11
// cell-iterations/s with the ordering below, profiles/k1_sweep3_rcm_r01.jsonl). Cuthill-McKee (breadth-first levels of the
// symmetrised graph, neighbours by ascending degree, started from a pseudo-peripheral cell) puts graph neighbours
// close together, so the rows gathered by one wave of CTAs form a narrow window of the permuted iterate.
//
// A second pass (grow_groups) regroups that order into tight sets of 32 cells that share sources: +3% on the gather
// (L1 hits between the rows of one CTA) at no run-time cost. Host cost at 50k cells / 1.5 M entries: Cuthill-McKee
// 46 ms, grouping 68 ms, the two span measurements 15 ms.
//
// Built once per plan on the host (O(nnz log k)); rejected when it does not shrink the mean |pos(i) - pos(j)| of the
// stored entries by at least 20% (an input that is already well ordered keeps its order).
#include <algorithm>
#include <numeric>
#include <stdlib.h>
#include <thread>

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
    auto bfs = [&](int32_t start, std::vector<int32_t> *out, std::vector<uint8_t> &mark, bool by_degree) -> int32_t {
        // returns the last node reached; appends the visit order to *out when given. by_degree: neighbours in
        // ascending degree (the Cuthill-McKee rule)
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
            if (by_degree)
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
        const int32_t far = bfs(v, &comp, tmp, true);
        for (int32_t u : comp) tmp[u] = 0;
        bfs(far, &order, seen, true);
    }
}

// mean |pos(i) - pos(j)| over the stored entries of every 8th destination (a sample is enough for the 20% test)
static double mean_span(int32_t n, const std::vector<int32_t> &in_ptr, const std::vector<Pair> &in_pairs,
                        const int32_t *pos) {
    double s = 0.0;
    size_t cnt = 0;
    for (int32_t j = 0; j < n; j += 8)
        for (int32_t p = in_ptr[j]; p < in_ptr[j + 1]; ++p, ++cnt) {
            const int64_t a = pos ? pos[j] : j, b = pos ? pos[in_pairs[p].src] : in_pairs[p].src;
            s += (double)(a > b ? a - b : b - a);
        }
    return cnt == 0 ? 0.0 : s / (double)cnt;
}

// Second pass over the Cuthill-McKee order: grow tight groups of `group` cells. Seeds are taken in Cuthill-McKee order
// (so the groups still sweep the graph as a narrow band and the L2 window is kept); a group is grown by repeatedly
// taking the unplaced cell with the most neighbours already in the group. Cells of one group share many sources,
// which is what the staged gather (one group = one CTA's shared memory) turns into saved L2 -> SM traffic:
// stored entries per distinct source of 32 consecutive rows, 50k cells / k = 30: 1.01 caller order, 1.57 CM, 1.98 here.
static void grow_groups(int32_t n, int32_t group, const std::vector<int32_t> &out_ptr, const std::vector<int32_t> &out_idx,
                        const std::vector<int32_t> &in_ptr, const std::vector<Pair> &in_pairs,
                        const std::vector<int32_t> &seed_order, std::vector<int32_t> &order) {
    std::vector<uint8_t> placed(n, 0);
    std::vector<int32_t> links(n, 0), touched;
    // bucket[c]: cells pushed when their link count became c (entries go stale when the count moves on or the cell is
    // placed). The next member is the lowest id among the valid entries of the highest non-empty bucket.
    std::vector<std::vector<int32_t>> bucket;
    order.clear();
    order.reserve(n);
    size_t seed_ptr = 0;
    while ((int32_t)order.size() < n) {
        while (placed[seed_order[seed_ptr]]) ++seed_ptr;
        int32_t v = seed_order[seed_ptr];
        const size_t first = order.size();
        touched.clear();
        int32_t top = 0;
        for (int32_t got = 0; got < group; ++got) {
            placed[v] = 1;
            order.push_back(v);
            auto visit = [&](int32_t u) {
                if (placed[u]) return;
                const int32_t c = ++links[u];
                if (c == 1) touched.push_back(u);
                if ((size_t)c >= bucket.size()) bucket.resize((size_t)c + 1);
                bucket[c].push_back(u);
                if (c > top) top = c;
            };
            for (int32_t p = out_ptr[v]; p < out_ptr[v + 1]; ++p) visit(out_idx[p]);
            for (int32_t p = in_ptr[v]; p < in_ptr[v + 1]; ++p) visit(in_pairs[p].src);
            int32_t best = -1;
            while (top > 0) {
                std::vector<int32_t> &b = bucket[top];
                size_t w = 0;
                for (size_t r = 0; r < b.size(); ++r) {
                    const int32_t u = b[r];
                    if (placed[u] || links[u] != top) continue;  // stale
                    b[w++] = u;
                    if (best < 0 || u < best) best = u;
                }
                b.resize(w);
                if (best >= 0) break;
                --top;
            }
            if (best < 0) break;  // the component is exhausted: the next group starts from the next seed
            v = best;
        }
        for (int32_t u : touched) links[u] = 0;
        for (size_t q = first; q < order.size(); ++q) links[order[q]] = 0;
        for (std::vector<int32_t> &b : bucket) b.clear();  // at most max-degree buckets
    }
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
    std::vector<float> in_lo;  // residuals of a float64 matrix, entry for entry with in_pairs
    if (plan->d_t_lo) {
        in_lo.resize(nnz);
        CY2P_CUDA(cudaMemcpyAsync(in_lo.data(), plan->d_t_lo, sizeof(float) * nnz, cudaMemcpyDeviceToHost, st));
    }
    CY2P_CUDA(cudaStreamSynchronize(st));

    std::vector<int32_t> order;
    cuthill_mckee(n, out_ptr, out_idx, in_ptr, in_pairs, order);
    if ((int32_t)order.size() != n) return CY2P_OK;
    const char *gv = getenv("CY2P_GROUP");  // 0 keeps the plain Cuthill-McKee order
    const int group = (gv && *gv) ? atoi(gv) : 32;
    if (group > 1) {
        std::vector<int32_t> grouped;
        grow_groups(n, group, out_ptr, out_idx, in_ptr, in_pairs, order, grouped);
        if ((int32_t)grouped.size() == n) order.swap(grouped);
    }
    std::vector<int32_t> pos(n);
    for (int32_t q = 0; q < n; ++q) pos[order[q]] = q;
    plan->span_identity = mean_span(n, in_ptr, in_pairs, nullptr);
    plan->span_reordered = mean_span(n, in_ptr, in_pairs, pos.data());
    if (!(plan->span_reordered < 0.8 * plan->span_identity)) return CY2P_OK;  // already well ordered

    std::vector<int32_t> r_ptr((size_t)n + 1);
    std::vector<Pair> pp(nnz), op(nnz);
    std::vector<float> lo_pp(in_lo.size()), lo_op(in_lo.size());
    r_ptr[0] = 0;
    for (int32_t q = 0; q < n; ++q) r_ptr[q + 1] = r_ptr[q] + (in_ptr[order[q] + 1] - in_ptr[order[q]]);
    // rows are independent: a few host threads share the permute + per-row sort (95 -> ~25 ms at 1.5 M entries)
    auto build_rows = [&](int32_t q0, int32_t q1) {
        for (int32_t q = q0; q < q1; ++q) {
            const int32_t j = order[q];
            const int32_t len = in_ptr[j + 1] - in_ptr[j];
            Pair *dpp = pp.data() + r_ptr[q], *dop = op.data() + r_ptr[q];
            for (int32_t t = 0; t < len; ++t) {
                const Pair e = in_pairs[in_ptr[j] + t];
                dop[t] = e;  // ascending original source, as in T^T
                dpp[t] = Pair{pos[e.src], e.val};
            }
            if (in_lo.empty()) {
                std::sort(dpp, dpp + len, [](const Pair &a, const Pair &b) { return a.src < b.src; });  // ascending addresses
                continue;
            }
            struct Wide { Pair p; float lo; };
            std::vector<Wide> w((size_t)len);
            for (int32_t t = 0; t < len; ++t) {
                lo_op[r_ptr[q] + t] = in_lo[in_ptr[j] + t];
                w[t] = Wide{dpp[t], in_lo[in_ptr[j] + t]};
            }
            std::sort(w.begin(), w.end(), [](const Wide &a, const Wide &b) { return a.p.src < b.p.src; });
            for (int32_t t = 0; t < len; ++t) {
                dpp[t] = w[t].p;
                lo_pp[r_ptr[q] + t] = w[t].lo;
            }
        }
    };
    {
        unsigned nthreads = std::thread::hardware_concurrency();
        nthreads = nthreads == 0 ? 1 : (nthreads > 8 ? 8 : nthreads);
        if (n < 20000) nthreads = 1;
        std::vector<std::thread> pool;
        const int32_t chunk = (n + (int32_t)nthreads - 1) / (int32_t)nthreads;
        for (unsigned t = 1; t < nthreads; ++t)
            pool.emplace_back(build_rows, std::min(n, (int32_t)t * chunk), std::min(n, (int32_t)(t + 1) * chunk));
        build_rows(0, std::min(n, chunk));
        for (std::thread &th : pool) th.join();
    }
    // Staged-gather blocks: consecutive permuted rows while their distinct sources fit `ucap` shared-memory segments.
    const char *uv = getenv("CY2P_STAGE_UCAP");
    const int ucap = (uv && *uv) ? atoi(uv) : 272;
    const char *rv = getenv("CY2P_STAGE_ROWS");
    const int rmax = std::min(64, (rv && *rv) ? atoi(rv) : 32);  // spmm_staged: 8 warps x at most 8 rows each
    std::vector<int32_t> blk_row, blk_uptr, blk_usrc;
    std::vector<Pair> lp;
    const char *sv = getenv("CY2P_STAGED");  // the staged gather is an opt-in experiment (propagate.cu)
    const bool want_blocks = sv && *sv && atoi(sv) != 0;
    if (want_blocks && ucap >= plan->max_in_deg && plan->max_in_deg <= 2048 && ucap > 0 && rmax > 0) {
        lp.resize(nnz);
        std::vector<int32_t> stamp(n, -1), local(n, 0), cur;
        blk_row.push_back(0);
        blk_uptr.push_back(0);
        int32_t q = 0;
        while (q < n) {
            const int32_t b = (int32_t)blk_row.size() - 1;
            cur.clear();
            int32_t q1 = q;
            while (q1 < n && q1 - q < rmax) {
                if (q1 > q && r_ptr[q1 + 1] - r_ptr[q] > 2048) break;  // spmm_staged keeps a block's pairs in shared memory
                size_t added = 0;
                for (int32_t p = r_ptr[q1]; p < r_ptr[q1 + 1]; ++p)
                    if (stamp[pp[p].src] != b) {
                        stamp[pp[p].src] = b;
                        cur.push_back(pp[p].src);
                        ++added;
                    }
                if ((int32_t)cur.size() > ucap && q1 > q) {  // this row does not fit any more: undo it
                    for (size_t t = 0; t < added; ++t) {
                        stamp[cur.back()] = -1;
                        cur.pop_back();
                    }
                    break;
                }
                ++q1;
            }
            std::sort(cur.begin(), cur.end());
            for (size_t t = 0; t < cur.size(); ++t) local[cur[t]] = (int32_t)t;
            for (int32_t p = r_ptr[q]; p < r_ptr[q1]; ++p) lp[p] = Pair{local[pp[p].src], pp[p].val};
            blk_usrc.insert(blk_usrc.end(), cur.begin(), cur.end());
            blk_uptr.push_back((int32_t)blk_usrc.size());
            blk_row.push_back(q1);
            q = q1;
        }
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
    if (!in_lo.empty()) {
        if (e == cudaSuccess) e = upload((void **)&plan->d_r_lo_pp, lo_pp.data(), sizeof(float) * nnz);
        if (e == cudaSuccess) e = upload((void **)&plan->d_r_lo_op, lo_op.data(), sizeof(float) * nnz);
    }
    if (!blk_usrc.empty()) {
        if (e == cudaSuccess) e = upload((void **)&plan->d_blk_row, blk_row.data(), sizeof(int32_t) * blk_row.size());
        if (e == cudaSuccess) e = upload((void **)&plan->d_blk_uptr, blk_uptr.data(), sizeof(int32_t) * blk_uptr.size());
        if (e == cudaSuccess) e = upload((void **)&plan->d_blk_usrc, blk_usrc.data(), sizeof(int32_t) * blk_usrc.size());
        if (e == cudaSuccess) e = upload((void **)&plan->d_r_pairs_lp, lp.data(), sizeof(Pair) * nnz);
        if (e == cudaSuccess) {
            plan->nblk = (int32_t)blk_row.size() - 1;
            plan->blk_ucap = ucap;
            plan->blk_reuse = (double)nnz / (double)blk_usrc.size();
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the host vectors go out of scope
    if (e != cudaSuccess) {
        set_error("reorder: %s", cudaGetErrorString(e));
        return CY2P_ERR_CUDA;
    }
    plan->reorder_state = 1;
    return CY2P_OK;
}

}  // namespace cy2p

// Host-only entry point (no device needed): the locality order of a CSR matrix, the same passes ensure_reordered runs
// for a plan. The row-partitioned driver (cy2path_b200/rowpart.py) orders the WHOLE graph once on the host, cuts the
// order into per-rank blocks and builds each rank's plan from its slice only, so no rank ever holds the whole T on
// its device. order_out[q] = original cell id at position q. group <= 1 keeps the plain Cuthill-McKee order.
extern "C" int32_t cy2p_order_host(int32_t n, int64_t nnz, const int32_t *h_indptr, const int32_t *h_indices,
                                   int32_t group, int32_t *order_out) {
    using namespace cy2p;
    CY2P_REQUIRE(n >= 0 && nnz >= 0 && h_indptr && order_out && (nnz == 0 || h_indices), "null pointer / negative size");
    if (h_indptr[0] != 0 || h_indptr[n] != nnz) {
        set_error("indptr[0] must be 0 and indptr[n] == nnz");
        return CY2P_ERR_BAD_CSR;
    }
    std::vector<int32_t> out_ptr(h_indptr, h_indptr + n + 1), out_idx(h_indices, h_indices + nnz), in_ptr((size_t)n + 1, 0);
    for (int32_t i = 0; i < n; ++i)
        if (out_ptr[i + 1] < out_ptr[i]) {
            set_error("indptr is not monotone at row %d", i);
            return CY2P_ERR_BAD_CSR;
        }
    for (int64_t p = 0; p < nnz; ++p) {
        if (out_idx[p] < 0 || out_idx[p] >= n) {
            set_error("column index %d out of range at entry %lld", out_idx[p], (long long)p);
            return CY2P_ERR_BAD_CSR;
        }
        ++in_ptr[out_idx[p] + 1];
    }
    for (int32_t j = 0; j < n; ++j) in_ptr[j + 1] += in_ptr[j];
    std::vector<Pair> in_pairs((size_t)nnz);
    {
        std::vector<int32_t> cur(in_ptr.begin(), in_ptr.end() - 1);
        for (int32_t i = 0; i < n; ++i)
            for (int32_t p = out_ptr[i]; p < out_ptr[i + 1]; ++p) in_pairs[cur[out_idx[p]]++] = Pair{i, 0.f};
    }
    std::vector<int32_t> order;
    cuthill_mckee(n, out_ptr, out_idx, in_ptr, in_pairs, order);
    if ((int32_t)order.size() != n) {
        set_error("ordering failed");
        return CY2P_ERR_INVALID_ARG;
    }
    if (group > 1) {
        std::vector<int32_t> grouped;
        grow_groups(n, group, out_ptr, out_idx, in_ptr, in_pairs, order, grouped);
        if ((int32_t)grouped.size() == n) order.swap(grouped);
    }
    std::copy(order.begin(), order.end(), order_out);
    return CY2P_OK;
}

