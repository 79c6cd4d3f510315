// Host-only lab for the locality ordering of the batched gather (no GPU needed).
//
// Includes reorder.cu itself (its ordering passes are plain host C++), runs the plan's own order and candidate
// orders on a CSR dump of T and prints the proxies the gather kernel's L1 / L2 behaviour follows:
//   share(R)   stored entries per distinct source of R consecutive destination rows (R = 8: the rows the 8 warps of a
//              CTA form at the same time; R = 32 / 64: a group / a CTA)
//   span       mean |pos(dest) - pos(src)| over the stored entries (width of the window of X a wave of CTAs reads)
//   lru(C)     hit rate of an LRU of C 512-byte row segments under the access stream of one CTA (64 rows, 8 warps,
//              a warp walks its row's sources four at a time, warps interleaved) — a stand-in for the share of the
//              SM's L1 one of its six resident CTAs gets
//
//   g++ -O2 -std=c++17 -I/usr/local/cuda/include tools/reorder_lab.cpp -o /tmp/reorder_lab -L/usr/local/cuda/lib64 -lcudart
//   /tmp/reorder_lab graph.bin      (graph.bin: int64 n, int64 nnz, int32 indptr[n+1], int32 indices[nnz], float data[nnz])
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../cy2path_b200/csrc/reorder.cu"

namespace cy2p {
void set_error(const char *, ...) {}
}  // namespace cy2p

using namespace cy2p;
using Clock = std::chrono::steady_clock;

struct Graph {
    int32_t n;
    std::vector<int32_t> out_ptr, out_idx, in_ptr;
    std::vector<Pair> in_pairs;
};

static Graph load(const char *path) {
    Graph g;
    FILE *f = fopen(path, "rb");
    if (!f) {
        perror(path);
        exit(1);
    }
    int64_t hdr[2];
    if (fread(hdr, 8, 2, f) != 2) exit(1);
    g.n = (int32_t)hdr[0];
    const size_t nnz = (size_t)hdr[1];
    g.out_ptr.resize(g.n + 1);
    g.out_idx.resize(nnz);
    std::vector<float> data(nnz);
    if (fread(g.out_ptr.data(), 4, g.n + 1, f) != (size_t)g.n + 1 || fread(g.out_idx.data(), 4, nnz, f) != nnz ||
        fread(data.data(), 4, nnz, f) != nnz)
        exit(1);
    fclose(f);
    // T^T with ascending sources per destination (what plan.cu's stable sort produces)
    g.in_ptr.assign(g.n + 1, 0);
    for (size_t p = 0; p < nnz; ++p) ++g.in_ptr[g.out_idx[p] + 1];
    for (int32_t j = 0; j < g.n; ++j) g.in_ptr[j + 1] += g.in_ptr[j];
    g.in_pairs.resize(nnz);
    std::vector<int32_t> cur(g.in_ptr.begin(), g.in_ptr.end() - 1);
    for (int32_t i = 0; i < g.n; ++i)
        for (int32_t p = g.out_ptr[i]; p < g.out_ptr[i + 1]; ++p) g.in_pairs[cur[g.out_idx[p]]++] = Pair{i, data[p]};
    return g;
}

// ------------------------------------------------------------------------------------------------ candidates
// Greedy groups by SHARED SOURCES: a group's next member is the unplaced cell with the most T^T entries whose source
// is already in the group's source set (ties: earliest in the seed order, which keeps the band narrow). The plan's
// grow_groups counts graph neighbours that are group members instead.
static void grow_groups_shared(const Graph &g, int32_t group, const std::vector<int32_t> &seed_order,
                               std::vector<int32_t> &order, int norm) {
    const int32_t n = g.n;
    std::vector<int32_t> seed_pos(n);
    for (int32_t q = 0; q < n; ++q) seed_pos[seed_order[q]] = q;
    std::vector<uint8_t> placed(n, 0);
    std::vector<int32_t> score(n, 0), src_stamp(n, -1), touched;
    std::vector<std::vector<int32_t>> bucket;
    order.clear();
    order.reserve(n);
    size_t seed_ptr = 0;
    int32_t gid = 0;
    while ((int32_t)order.size() < n) {
        while (placed[seed_order[seed_ptr]]) ++seed_ptr;
        int32_t v = seed_order[seed_ptr];
        touched.clear();
        int32_t top = 0;
        for (int32_t got = 0; got < group; ++got) {
            placed[v] = 1;
            order.push_back(v);
            for (int32_t p = g.in_ptr[v]; p < g.in_ptr[v + 1]; ++p) {
                const int32_t s = g.in_pairs[p].src;
                if (src_stamp[s] == gid) continue;
                src_stamp[s] = gid;  // a new source of the group: every other destination of s now shares it
                for (int32_t e = g.out_ptr[s]; e < g.out_ptr[s + 1]; ++e) {
                    const int32_t u = g.out_idx[e];
                    if (placed[u]) continue;
                    const int32_t c = ++score[u];
                    if (c == 1) touched.push_back(u);
                    if ((size_t)c >= bucket.size()) bucket.resize((size_t)c + 1);
                    bucket[c].push_back(u);
                    if (c > top) top = c;
                }
            }
            int32_t best = -1;
            if (norm == 0) {
                while (top > 0) {
                    std::vector<int32_t> &b = bucket[top];
                    size_t w = 0;
                    for (size_t r = 0; r < b.size(); ++r) {
                        const int32_t u = b[r];
                        if (placed[u] || score[u] != top) continue;
                        b[w++] = u;
                        if (best < 0 || seed_pos[u] < seed_pos[best]) best = u;
                    }
                    b.resize(w);
                    if (best >= 0) break;
                    --top;
                }
            } else {  // fewest NEW sources (in-degree - shared), ties by shared count then seed position
                int32_t best_new = 1 << 30;
                for (int32_t u : touched) {
                    if (placed[u]) continue;
                    const int32_t nw = (g.in_ptr[u + 1] - g.in_ptr[u]) - score[u];
                    if (nw < best_new || (nw == best_new && (score[u] > score[best] ||
                                                             (score[u] == score[best] && seed_pos[u] < seed_pos[best])))) {
                        best = u;
                        best_new = nw;
                    }
                }
            }
            if (best < 0) break;
            v = best;
        }
        for (int32_t u : touched) score[u] = 0;
        for (std::vector<int32_t> &b : bucket) b.clear();
        ++gid;
    }
}

// ------------------------------------------------------------------------------------------------ proxies
struct Perm {
    std::vector<int32_t> r_ptr;
    std::vector<int32_t> src;  // permuted source positions, ascending inside a row
};

static Perm permute(const Graph &g, const std::vector<int32_t> &order) {
    const int32_t n = g.n;
    std::vector<int32_t> pos(n);
    for (int32_t q = 0; q < n; ++q) pos[order[q]] = q;
    Perm P;
    P.r_ptr.assign(n + 1, 0);
    for (int32_t q = 0; q < n; ++q) P.r_ptr[q + 1] = P.r_ptr[q] + (g.in_ptr[order[q] + 1] - g.in_ptr[order[q]]);
    P.src.resize(g.in_pairs.size());
    for (int32_t q = 0; q < n; ++q) {
        const int32_t j = order[q];
        int32_t *d = P.src.data() + P.r_ptr[q];
        const int32_t len = g.in_ptr[j + 1] - g.in_ptr[j];
        for (int32_t t = 0; t < len; ++t) d[t] = pos[g.in_pairs[g.in_ptr[j] + t].src];
        std::sort(d, d + len);
    }
    return P;
}

static double share(const Perm &P, int32_t n, int32_t R) {
    std::vector<int32_t> stamp(n, -1);
    size_t refs = 0, distinct = 0;
    for (int32_t q0 = 0, b = 0; q0 < n; q0 += R, ++b) {
        const int32_t q1 = std::min(n, q0 + R);
        for (int32_t p = P.r_ptr[q0]; p < P.r_ptr[q1]; ++p) {
            ++refs;
            if (stamp[P.src[p]] != b) {
                stamp[P.src[p]] = b;
                ++distinct;
            }
        }
    }
    return (double)refs / (double)distinct;
}

static double span(const Perm &P, int32_t n) {
    double s = 0;
    for (int32_t q = 0; q < n; ++q)
        for (int32_t p = P.r_ptr[q]; p < P.r_ptr[q + 1]; ++p) s += std::abs((double)P.src[p] - (double)q);
    return s / (double)P.src.size();
}

// LRU of `cap` segments under the access stream of one CTA: rows_per_cta rows, 8 warps, warp w forms rows
// q0 + w, q0 + w + 8, ...; the warps take turns, four gathers per turn.
static double lru_hit(const Perm &P, int32_t n, int32_t rows_per_cta, int32_t cap) {
    constexpr int W = 8;
    std::vector<int32_t> tag(cap);
    std::vector<int64_t> when(cap);
    size_t hits = 0, total = 0;
    for (int32_t q0 = 0; q0 < n; q0 += rows_per_cta) {
        const int32_t q1 = std::min(n, q0 + rows_per_cta);
        std::fill(tag.begin(), tag.end(), -1);
        std::fill(when.begin(), when.end(), -1);
        int64_t clock = 0;
        int32_t row[W], cur[W];
        for (int w = 0; w < W; ++w) {
            row[w] = q0 + w;
            cur[w] = row[w] < q1 ? P.r_ptr[row[w]] : 0;
        }
        bool live = true;
        while (live) {
            live = false;
            for (int w = 0; w < W; ++w) {
                if (row[w] >= q1) continue;
                live = true;
                for (int k = 0; k < 4 && cur[w] < P.r_ptr[row[w] + 1]; ++k, ++cur[w]) {
                    const int32_t s = P.src[cur[w]];
                    ++total;
                    int slot = -1, oldest = 0;
                    for (int c = 0; c < cap; ++c) {
                        if (tag[c] == s) {
                            slot = c;
                            break;
                        }
                        if (when[c] < when[oldest]) oldest = c;
                    }
                    if (slot >= 0)
                        ++hits;
                    else {
                        slot = oldest;
                        tag[slot] = s;
                    }
                    when[slot] = clock++;
                }
                if (cur[w] >= P.r_ptr[row[w] + 1]) {
                    row[w] += W;
                    if (row[w] < q1) cur[w] = P.r_ptr[row[w]];
                }
            }
        }
    }
    return (double)hits / (double)total;
}

static void report(const char *name, const Graph &g, const std::vector<int32_t> &order, double ms) {
    const Perm P = permute(g, order);
    printf("%-28s %7.0f ms  share8 %.3f share32 %.3f share64 %.3f  span %8.0f  lru64 %.3f lru128 %.3f lru256 %.3f\n", name,
           ms, share(P, g.n, 8), share(P, g.n, 32), share(P, g.n, 64), span(P, g.n), lru_hit(P, g.n, 64, 64),
           lru_hit(P, g.n, 64, 128), lru_hit(P, g.n, 64, 256));
    fflush(stdout);
}

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s graph.bin\n", argv[0]);
        return 2;
    }
    const Graph g = load(argv[1]);
    printf("n = %d, nnz = %zu\n", g.n, g.in_pairs.size());
    auto ms_since = [](Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); };
    std::vector<int32_t> ident(g.n);
    std::iota(ident.begin(), ident.end(), 0);
    report("caller order", g, ident, 0);
    std::vector<int32_t> cm;
    auto t0 = Clock::now();
    cuthill_mckee(g.n, g.out_ptr, g.out_idx, g.in_ptr, g.in_pairs, cm);
    const double cm_ms = ms_since(t0);
    report("cuthill-mckee", g, cm, cm_ms);
    for (int group : {32, 64}) {
        std::vector<int32_t> o;
        t0 = Clock::now();
        grow_groups(g.n, group, g.out_ptr, g.out_idx, g.in_ptr, g.in_pairs, cm, o);
        char nm[64];
        snprintf(nm, sizeof nm, "plan: CM + links g%d", group);
        report(nm, g, o, cm_ms + ms_since(t0));
    }
    for (int norm : {0, 1})
        for (int group : {32, 64, 128}) {
            std::vector<int32_t> o;
            t0 = Clock::now();
            grow_groups_shared(g, group, cm, o, norm);
            char nm[64];
            snprintf(nm, sizeof nm, "CM + shared-src%s g%d", norm ? "(new)" : "", group);
            report(nm, g, o, cm_ms + ms_since(t0));
        }
    return 0;
}
