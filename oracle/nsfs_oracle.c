/* TEST INFRASTRUCTURE ONLY — never linked into, imported by, or executed from the product path.
 *
 * Plain-C restatement of the reference planner hot path (mukundbala/Nav-Stack-From-Scratch,
 * src/global_planner/{include,src}/algo + src/bot_utils).  It restates WHAT the reference computes,
 * quirk for quirk, in flat-key form; every function cites the reference lines it
 * follows.  It is pinned two ways (tests/test_oracle_vs_reference.py, tests/golden/):
 *   1. against oracle/_ref/libnsfs_ref.so — the unmodified reference sources compiled here — on
 *      seeded maps/queries (paths, g-fields, fallback cells, Bresenham rays, post-processing);
 *   2. against the committed golden fixtures generated from that same reference build.
 * The reference repo itself holds no tests, fixtures or golden vectors.
 *
 * ThetaStar is the exception: the reference's thetastar.cpp is 0 bytes and thetastar.h:4-7 is an
 * empty class, so nso_plan(NSO_THETASTAR) is a DESIGN oracle (Basic Theta* built from the reference's
 * own primitives: OpenList, testPos, isLos/bresenham_los, dist_euc).  PARITY UNPINNED for that path.
 *
 * Third-party arithmetic on the path: libstdc++ std::priority_queue / push_heap / pop_heap
 * (bits/stl_heap.h __push_heap / __adjust_heap, GCC 13.3 here; same algorithm in the GCC 9 of
 * ros:noetic).  heap_push()/heap_pop() below restate that published algorithm; nso_heap_trace is
 * checked against the reference's own OpenList driven through the real std::priority_queue.
 */
#include "nsfs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_SQRT2
#define M_SQRT2 1.41421356237309504880
#endif

/* NB_LUT, grid_planner_core.h:197 */
static const int DI[8] = {1, 1, 0, -1, -1, -1, 0, 1};
static const int DJ[8] = {0, 1, 1, 1, 0, -1, -1, -1};

#define G_INF 1e5       /* astar.cpp:40, djikstra.cpp:40,175 */
#define G_SLACK 1e-5    /* astar.cpp:116, djikstra.cpp:113,258 */

/* ------------------------------------------------------------------ OpenList */

/* OpenNode, grid_planner_core.h:35-46: the constructors take int f_arg, int g_arg, so the doubles
 * handed over by pushToOpenList (grid_planner_core.cpp:385-391) are truncated toward zero. */
typedef struct { int f, g, key; } open_node;

typedef struct { open_node* a; long long n, cap; int mode; } open_list;

/* comparators, grid_planner_core.cpp:55-75 ("a sorts after b") */
static int ol_comp(int mode, const open_node* a, const open_node* b) {
    if (mode == NSO_MODE_F) return a->f > b->f;
    if (mode == NSO_MODE_G) return a->g > b->g;
    if (a->f == b->f) return a->g > b->g;
    return a->f > b->f;
}

/* std::priority_queue::push = push_back + std::push_heap -> __push_heap (stl_heap.h:135-148) */
static int heap_push(open_list* q, open_node v) {
    if (q->n == q->cap) {
        long long nc = q->cap ? q->cap * 2 : 1024;
        open_node* na = (open_node*)realloc(q->a, (size_t)nc * sizeof(open_node));
        if (!na) return NSO_E_NOMEM;
        q->a = na; q->cap = nc;
    }
    long long hole = q->n++;
    while (hole > 0) {
        long long parent = (hole - 1) / 2;
        if (!ol_comp(q->mode, &q->a[parent], &v)) break;
        q->a[hole] = q->a[parent];
        hole = parent;
    }
    q->a[hole] = v;
    return NSO_OK;
}

/* top() + pop(): std::pop_heap -> __pop_heap -> __adjust_heap (stl_heap.h:224-262), then pop_back.
 * The hole always sinks to a leaf (preferring the right child unless it sorts after the left one),
 * then the displaced last element is sifted up from there. */
static open_node heap_pop(open_list* q) {
    open_node top = q->a[0];
    long long len = q->n - 1;            /* length after removing the last slot */
    if (len > 0) {
        open_node v = q->a[len];
        long long hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (ol_comp(q->mode, &q->a[child], &q->a[child - 1])) child--;
            q->a[hole] = q->a[child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            q->a[hole] = q->a[child - 1];
            hole = child - 1;
        }
        while (hole > 0) {
            long long parent = (hole - 1) / 2;
            if (!ol_comp(q->mode, &q->a[parent], &v)) break;
            q->a[hole] = q->a[parent];
            hole = parent;
        }
        q->a[hole] = v;
    }
    q->n = len;
    return top;
}

int nso_heap_trace(int mode, int n_ops, const int32_t* ops, int32_t* out) {
    open_list q = {0, 0, 0, mode};
    int np = 0;
    for (int t = 0; t < n_ops; ++t) {
        const int32_t* o = ops + 4 * t;
        if (o[0] == 1) {
            open_node v = {o[1], o[2], o[3]};
            heap_push(&q, v);
        } else if (q.n > 0) {
            out[np++] = heap_pop(&q).key;
        }
    }
    free(q.a);
    return np;
}

/* ------------------------------------------------------------------ bot_utils */

/* dist_oct, bot_utils.cpp:44-65.  Line 47 subtracts src_x where src_y is meant, so for Index
 * arguments the result is oct(|ti - si|, |tj - si|): it ignores the source column. */
double nso_dist_oct(int si, int sj, int ti, int tj) {
    (void)sj;
    double ax = fabs((double)ti - (double)si);
    double ay = fabs((double)tj - (double)si);
    double ord, card;
    if (ax > ay) { ord = ay; card = ax - ay; } else { ord = ax; card = ay - ax; }
    return ord * M_SQRT2 + card;
}

/* dist_euc, bot_utils.cpp:72-87 */
double nso_dist_euc(int si, int sj, int ti, int tj) {
    double dx = (double)ti - (double)si, dy = (double)tj - (double)sj;
    return sqrt(dx * dx + dy * dy);
}

/* idx2pos / pos2idx, grid_planner_core.cpp:409-421 */
void nso_idx2pos(const nso_map* m, int i, int j, double* x, double* y) {
    *x = i * m->cell + m->ox;
    *y = j * m->cell + m->oy;
}
void nso_pos2idx(const nso_map* m, double x, double y, int* i, int* j) {
    *i = (int)round((x - m->ox) / m->cell);
    *j = (int)round((y - m->oy) / m->cell);
}

static int isign(int v) { return (v > 0) - (v < 0); }   /* bot_utils::sign, bot_utils.cpp:34-42 */

/* bresenham_los, bot_utils.cpp:137-192 (iterative form, as the reference runs it) */
int nso_bresenham(int si, int sj, int ti, int tj, int32_t* out, int cap) {
    int Di = ti - si, Dj = tj - sj;
    int Dl, Ds, l, s, lf, sf, long_is_i;
    if (abs(Di) > abs(Dj)) { Dl = Di; Ds = Dj; l = si; s = sj; lf = ti; sf = tj; long_is_i = 1; }
    else                   { Dl = Dj; Ds = Di; l = sj; s = si; lf = tj; sf = ti; long_is_i = 0; }
    int ds = isign(Ds), dl = isign(Dl), abs_Dl = abs(Dl), d_esl = abs(Dl) * ds, esl = 0;
    int n = 0;
    if (n < cap) { out[0] = si; out[1] = sj; }
    n++;
    while (l != lf || s != sf) {
        l += dl;
        esl += Ds;
        if (2 * abs(esl) >= abs_Dl) { esl -= d_esl; s += ds; }
        if (n < cap) { out[2 * n] = long_is_i ? l : s; out[2 * n + 1] = long_is_i ? s : l; }
        n++;
    }
    return n;
}

/* Closed form of the same ray: step k has l = l0 + k*dl and
 * s = s0 + ds * floor((2k|Ds| + |Dl|) / (2|Dl|)); this is what the CUDA line-of-sight kernel evaluates
 * one lane per cell.  tests/ check it cell-for-cell against nso_bresenham and the reference. */
int nso_bresenham_closed(int si, int sj, int ti, int tj, int32_t* out, int cap) {
    int Di = ti - si, Dj = tj - sj;
    int long_is_i = abs(Di) > abs(Dj);
    int Dl = long_is_i ? Di : Dj, Ds = long_is_i ? Dj : Di;
    int l0 = long_is_i ? si : sj, s0 = long_is_i ? sj : si;
    int aDl = abs(Dl), aDs = abs(Ds), dl = isign(Dl), ds = isign(Ds);
    int n = aDl + 1;
    for (int k = 0; k < n && k < cap; ++k) {
        int l = l0 + k * dl;
        int s = s0 + (aDl ? ds * (int)((2LL * k * aDs + aDl) / (2LL * aDl)) : 0);
        out[2 * k] = long_is_i ? l : s;
        out[2 * k + 1] = long_is_i ? s : l;
    }
    return n;
}

/* ------------------------------------------------------------------ GridPlannerCore helpers */

/* testPos(Index), grid_planner_core.cpp:434-455 with oob(Index) from 423-426.  oob only ever tests
 * the row ("idx.j < 0 && idx.j >= W" is always false), and the key is formed before the test, so a
 * neighbour is addressed by the FLAT key i*W + j: column overflow wraps into the adjacent row, and a
 * key outside [0, N) with an in-range row makes std::vector::at throw std::out_of_range. */
int nso_test_pos(const nso_map* m, int i, int j) {
    long long key = (long long)i * m->W + j;
    if (i < 0 || i >= m->H) return 0;
    if (key < 0 || key >= (long long)m->H * m->W) return NSO_E_RANGE;
    if (m->infl[key] > 0) return 0;
    if (m->lo[key] > m->lo_thresh) return 0;
    return 1;
}

/* isLos(Index), grid_planner_core.cpp:481-496 */
int nso_is_los(const nso_map* m, int si, int sj, int ti, int tj) {
    int Di = abs(ti - si), Dj = abs(tj - sj);
    int n = (Di > Dj ? Di : Dj) + 1;
    int32_t* ray = (int32_t*)malloc((size_t)n * 2 * sizeof(int32_t));
    if (!ray) return NSO_E_NOMEM;
    nso_bresenham(si, sj, ti, tj, ray, n);
    int res = 1;
    for (int k = 0; k < n; ++k) {
        int t = nso_test_pos(m, ray[2 * k], ray[2 * k + 1]);
        if (t != 1) { res = t; break; }
    }
    free(ray);
    return res;
}

/* sparsifyPath, grid_planner_core.cpp:506-535 (abs resolves to the double overload, see stub/ros/ros.h) */
int nso_sparsify(const double* p, int n, double* out, int cap) {
    int m = 0;
#define EMIT(t) do { if (m < cap) { out[2 * m] = p[2 * (t)]; out[2 * m + 1] = p[2 * (t) + 1]; } m++; } while (0)
    if (n <= 2) { for (int t = 0; t < n; ++t) EMIT(t); return m; }
    EMIT(0);
    for (int t = 2; t < n; ++t) {
        double dxn = p[2 * t] - p[2 * (t - 1)], dyn = p[2 * t + 1] - p[2 * (t - 1) + 1];
        double dxp = p[2 * (t - 1)] - p[2 * (t - 2)], dyp = p[2 * (t - 1) + 1] - p[2 * (t - 2) + 1];
        if (fabs(dxn * dyp - dyn * dxp) > 1e-5) EMIT(t - 1);
    }
    EMIT(n - 1);
#undef EMIT
    return m;
}

/* makeAnyAnglePath, grid_planner_core.cpp:537-558 */
int nso_any_angle(const nso_map* m, const double* p, int n, double* out, int cap) {
    int k = 0;
#define EMITP(x, y) do { if (k < cap) { out[2 * k] = (x); out[2 * k + 1] = (y); } k++; } while (0)
    if (n < 1) return NSO_E_ARG;
    EMITP(p[0], p[1]);
    double cx = p[0], cy = p[1];
    int ci, cj;
    nso_pos2idx(m, cx, cy, &ci, &cj);
    for (int t = 1; t < n - 1; ++t) {
        int ti, tj;
        nso_pos2idx(m, p[2 * (t + 1)], p[2 * (t + 1) + 1], &ti, &tj);
        int los = nso_is_los(m, ci, cj, ti, tj);
        if (los < 0) return los;
        if (!los) {
            EMITP(p[2 * t], p[2 * t + 1]);
            nso_pos2idx(m, p[2 * t], p[2 * t + 1], &ci, &cj);
        }
    }
    EMITP(p[2 * (n - 1)], p[2 * (n - 1) + 1]);
#undef EMITP
    return k;
}

/* post_process_path, astar.cpp:150-165 == djikstra.cpp:147-163 */
int nso_post_process(const nso_map* m, const double* p, int n, double* out, int cap) {
    if (n <= 2) {
        for (int t = 0; t < n && t < cap; ++t) { out[2 * t] = p[2 * t]; out[2 * t + 1] = p[2 * t + 1]; }
        return n;
    }
    double* sp = (double*)malloc((size_t)n * 2 * sizeof(double));
    if (!sp) return NSO_E_NOMEM;
    int ns = nso_sparsify(p, n, sp, n);
    int r = nso_any_angle(m, sp, ns, out, cap);
    free(sp);
    return r;
}

/* ------------------------------------------------------------------ node store */

/* GridPlannerCore::Node (grid_planner_core.h:26-33) as three arrays plus a lazily-materialised
 * stamp (bit0 = reset this query, bit1 = visited) so the O(N) reset of astar.cpp:37-42 costs only
 * the pages a query touches; an unstamped cell reads as g = 1e5, visited = false. */
typedef struct { double* g; int32_t* parent; uint8_t* stamp; long long N; } node_store;

static int ns_alloc(node_store* s, long long N) {
    s->N = N;
    s->g = (double*)malloc((size_t)N * sizeof(double));
    s->parent = (int32_t*)malloc((size_t)N * sizeof(int32_t));
    s->stamp = (uint8_t*)calloc((size_t)N, 1);
    return (s->g && s->parent && s->stamp) ? NSO_OK : NSO_E_NOMEM;
}
static void ns_free(node_store* s) { free(s->g); free(s->parent); free(s->stamp); }
static double ns_g(const node_store* s, long long k) { return (s->stamp[k] & 1) ? s->g[k] : G_INF; }
static void ns_set(node_store* s, long long k, double g, int32_t parent) {
    s->g[k] = g; s->parent[k] = parent; s->stamp[k] |= 1;
}

/* pushToOpenList, grid_planner_core.cpp:385-391: f = (mode == "g") ? 0 : g + h; both truncated to int */
static int push_cell(open_list* q, int mode, long long key, double g, double h, nso_stats* st) {
    double f = (mode == NSO_MODE_G) ? 0 : g + h;
    open_node v = {(int)f, (int)g, (int)key};
    st->pushes++;
    int r = heap_push(q, v);
    if (q->n > st->heap_peak) st->heap_peak = q->n;
    return r;
}

static double heuristic(int algo, int i, int j, int gi, int gj) {
    if (algo == NSO_ASTAR) return nso_dist_oct(i, j, gi, gj);       /* astar.cpp:39 */
    if (algo == NSO_THETASTAR) return nso_dist_euc(i, j, gi, gj);   /* design (no reference) */
    return 0;                                                       /* djikstra.cpp:39 */
}

/* Astar::plan astar.cpp:27-148 and Djikstra::plan djikstra.cpp:27-145 (identical but for h).
 * NSO_THETASTAR swaps the relaxation for Basic Theta*'s two paths (design oracle). */
int nso_plan(const nso_map* m, int algo, int mode, int si, int sj, int gi, int gj,
             int32_t* path_ij, int cap, int* n_path, double* g_goal,
             double* g_out, int32_t* parent_out, uint8_t* visited_out, nso_stats* st_out) {
    const int H = m->H, W = m->W;
    const long long N = (long long)H * W;
    if (si < 0 || si >= H || sj < 0 || sj >= W) return NSO_E_ARG;  /* reference indexes nodes[] out of bounds: UB */
    if (mode < 0 || mode > 2) mode = NSO_MODE_FG;                   /* unknown mode => fg, grid_planner_core.cpp:274-281 */
    nso_stats st = {0, 0, 0, 0, 0};
    node_store ns;
    if (ns_alloc(&ns, N) != NSO_OK) { ns_free(&ns); return NSO_E_NOMEM; }
    open_list q = {0, 0, 0, mode};
    int status = NSO_OK, found = 0;
    const long long ks = (long long)si * W + sj;
    const long long kg = (gi >= 0 && gi < H && gj >= 0 && gj < W) ? (long long)gi * W + gj : -1;

    ns_set(&ns, ks, 0.0, (int32_t)ks);                              /* astar.cpp:45-52 */
    push_cell(&q, mode, ks, 0.0, heuristic(algo, si, sj, gi, gj), &st);

    while (q.n != 0) {                                              /* astar.cpp:54 */
        open_node top = heap_pop(&q);
        st.pops++;
        long long ku = top.key;
        if (ns.stamp[ku] & 2) continue;                             /* stale entry, astar.cpp:59-62 */
        ns.stamp[ku] |= 2;
        st.expansions++;
        int i = (int)(ku / W), j = (int)(ku % W);
        if (i == gi && j == gj) { found = 1; break; }               /* astar.cpp:68-85 */

        double gu = ns_g(&ns, ku);
        int card = 1;                                               /* is_cardinal, astar.cpp:88 */
        for (int d = 0; d < 8; ++d) {
            int t = nso_test_pos(m, i + DI[d], j + DJ[d]);          /* astar.cpp:97 */
            if (t == NSO_E_RANGE) { status = NSO_E_RANGE; goto done; }
            if (!t) continue;                                       /* toggle NOT flipped, astar.cpp:97-101 */
            long long kv = ku + (long long)DI[d] * W + DJ[d];
            double cand; int32_t par;
            if (algo == NSO_THETASTAR) {
                int vi = (int)(kv / W), vj = (int)(kv % W);
                long long kp = ns.parent[ku];
                int pi = (int)(kp / W), pj = (int)(kp % W);
                st.los_checks++;
                if (nso_is_los(m, pi, pj, vi, vj) == 1) { cand = ns_g(&ns, kp) + nso_dist_euc(pi, pj, vi, vj); par = (int32_t)kp; }
                else { cand = gu + nso_dist_euc(i, j, vi, vj); par = (int32_t)ku; }
            } else {
                cand = gu + (card ? 1 : M_SQRT2);                   /* astar.cpp:104-107 */
                par = (int32_t)ku;
            }
            if (ns_g(&ns, kv) > cand + G_SLACK) {                   /* astar.cpp:116-121: also for visited kv */
                ns_set(&ns, kv, cand, par);
                int vi = (int)(kv / W), vj = (int)(kv % W);
                if (push_cell(&q, mode, kv, cand, heuristic(algo, vi, vj, gi, gj), &st) != NSO_OK) { status = NSO_E_NOMEM; goto done; }
            }
            card = !card;                                           /* astar.cpp:123 */
        }
    }

done:
    if (status == NSO_OK) {
        int n = 0;
        if (found) {                                                /* back-track, astar.cpp:70-82 */
            long long k = kg;
            for (;;) {
                if (n < cap && path_ij) { path_ij[2 * n] = (int)(k / W); path_ij[2 * n + 1] = (int)(k % W); }
                n++;
                if (k == ks) break;
                k = ns.parent[k];
            }
        } else {                                                    /* astar.cpp:129-135 */
            if (n < cap && path_ij) { path_ij[0] = si; path_ij[1] = sj; }
            n = 1;
        }
        if (n_path) *n_path = n;
        if (n > cap && path_ij) status = NSO_E_TRUNC;
        if (g_goal) *g_goal = kg >= 0 ? ns_g(&ns, kg) : G_INF;
        for (long long k = 0; k < N && (g_out || parent_out || visited_out); ++k) {
            if (g_out) g_out[k] = ns_g(&ns, k);
            if (parent_out) parent_out[k] = ((ns.stamp[k] & 1) && k != ks) ? ns.parent[k] : -1;
            if (visited_out) visited_out[k] = (ns.stamp[k] >> 1) & 1;
        }
    }
    if (st_out) *st_out = st;
    free(q.a);
    ns_free(&ns);
    return status;
}

/* Djikstra::find_better_point, djikstra.cpp:165-273.  Runs on a planner prepared with cost mode "g"
 * (global_planner.cpp:127-128).  Non-free neighbours are traversable at +100 (inflated) / +1000
 * (occupied) with NO step cost; the cardinal toggle flips for every neighbour whose row is in range. */
int nso_find_better_point(const nso_map* m, int bi, int bj, int* fi, int* fj, nso_stats* st_out) {
    const int H = m->H, W = m->W, mode = NSO_MODE_G;
    const long long N = (long long)H * W;
    if (bi < 0 || bi >= H || bj < 0 || bj >= W) return NSO_E_ARG;
    nso_stats st = {0, 0, 0, 0, 0};
    node_store ns;
    if (ns_alloc(&ns, N) != NSO_OK) { ns_free(&ns); return NSO_E_NOMEM; }
    open_list q = {0, 0, 0, mode};
    int status = NSO_OK;
    const long long ks = (long long)bi * W + bj;
    *fi = bi; *fj = bj;                                             /* failure => idx2pos(idx_bad), djikstra.cpp:272 */
    ns_set(&ns, ks, 0.0, (int32_t)ks);
    push_cell(&q, mode, ks, 0.0, 0.0, &st);
    while (q.n != 0) {
        open_node top = heap_pop(&q);
        st.pops++;
        long long ku = top.key;
        if (ns.stamp[ku] & 2) continue;
        ns.stamp[ku] |= 2;
        st.expansions++;
        int i = (int)(ku / W), j = (int)(ku % W);
        if (nso_test_pos(m, i, j) == 1) { *fi = i; *fj = j; break; }   /* djikstra.cpp:205-211 */
        double gu = ns_g(&ns, ku);
        int card = 1;
        for (int d = 0; d < 8; ++d) {
            int ni = i + DI[d];
            if (ni < 0 || ni >= H) continue;                        /* oob(): row only, djikstra.cpp:222-226 */
            int t = nso_test_pos(m, ni, j + DJ[d]);                 /* djikstra.cpp:230 */
            if (t == NSO_E_RANGE) { status = NSO_E_RANGE; goto done; }
            long long kv = ku + (long long)DI[d] * W + DJ[d];
            double cand = gu;
            if (!t) {
                if (m->infl[kv] > 0) cand += 100;                   /* djikstra.cpp:235-239 */
                else if (m->lo[kv] > m->lo_thresh) cand += 1000;    /* djikstra.cpp:241-245 */
            } else {
                cand += card ? 1 : M_SQRT2;                         /* djikstra.cpp:250 */
            }
            if (ns_g(&ns, kv) > cand + G_SLACK) {                   /* djikstra.cpp:258-263 */
                ns_set(&ns, kv, cand, (int32_t)ku);
                if (push_cell(&q, mode, kv, cand, 0.0, &st) != NSO_OK) { status = NSO_E_NOMEM; goto done; }
            }
            card = !card;                                           /* djikstra.cpp:266 */
        }
    }
done:
    if (st_out) *st_out = st;
    free(q.a);
    ns_free(&ns);
    return status;
}

/* Label-correcting fixed point of the plan() edge rule, independent of any queue order: used by the
 * tests to show that the Djikstra g-field is order-independent.  FIFO worklist. */
int nso_field_fixed_point(const nso_map* m, int si, int sj, double* g) {
    const int H = m->H, W = m->W;
    const long long N = (long long)H * W;
    if (si < 0 || si >= H || sj < 0 || sj >= W) return NSO_E_ARG;
    int32_t* fifo = (int32_t*)malloc((size_t)N * sizeof(int32_t));
    uint8_t* inq = (uint8_t*)calloc((size_t)N, 1);
    if (!fifo || !inq) { free(fifo); free(inq); return NSO_E_NOMEM; }
    for (long long k = 0; k < N; ++k) g[k] = G_INF;
    long long head = 0, cnt = 0;
    long long ks = (long long)si * W + sj;
    g[ks] = 0; fifo[0] = (int32_t)ks; cnt = 1; inq[ks] = 1;
    int status = NSO_OK;
    while (cnt) {
        long long ku = fifo[head]; head = (head + 1) % N; cnt--; inq[ku] = 0;
        int i = (int)(ku / W), j = (int)(ku % W), card = 1;
        for (int d = 0; d < 8; ++d) {
            int t = nso_test_pos(m, i + DI[d], j + DJ[d]);
            if (t == NSO_E_RANGE) { status = NSO_E_RANGE; goto out; }
            if (!t) continue;
            long long kv = ku + (long long)DI[d] * W + DJ[d];
            double cand = g[ku] + (card ? 1 : M_SQRT2);
            if (g[kv] > cand + G_SLACK) {
                g[kv] = cand;
                if (!inq[kv]) { fifo[(head + cnt) % N] = (int32_t)kv; cnt++; inq[kv] = 1; }
            }
            card = !card;
        }
    }
out:
    free(fifo); free(inq);
    return status;
}

/* Warm-started variant of the fixed point above for ONE ROW-BLOCK SHARD of a partitioned map: g is
 * in/out, every finite cell is a source, and cells outside rows [row_lo, row_hi) take no in-edges (they exist only so
 * that their neighbours' direction costs see the true neighbourhood).  slack = the relaxation slack (the reference uses
 * 1e-5, astar.cpp:116; 0 makes the fixed point unique so sharded and monolithic solves can be compared bit for bit).
 * cap: candidates above it are not applied (the multi-GPU driver advances all shards band by band).
 * Returns the number of label improvements, or a negative NSO_E_* code. */
long long nso_field_relax_window(const nso_map* m, double* g, int row_lo, int row_hi, double slack, double cap) {
    const int H = m->H, W = m->W;
    const long long N = (long long)H * W;
    int32_t* fifo = (int32_t*)malloc((size_t)N * sizeof(int32_t));
    uint8_t* inq = (uint8_t*)calloc((size_t)N, 1);
    if (!fifo || !inq) { free(fifo); free(inq); return NSO_E_NOMEM; }
    long long head = 0, cnt = 0, improved = 0;
    for (long long k = 0; k < N; ++k)
        if (g[k] < G_INF) { fifo[cnt++] = (int32_t)k; inq[k] = 1; }
    long long status = 0;
    while (cnt) {
        long long ku = fifo[head]; head = (head + 1) % N; cnt--; inq[ku] = 0;
        int i = (int)(ku / W), j = (int)(ku % W), card = 1;
        for (int d = 0; d < 8; ++d) {
            int t = nso_test_pos(m, i + DI[d], j + DJ[d]);
            if (t == NSO_E_RANGE) {            /* would throw when expanded: only an error if this cell is a counted one */
                if (i >= row_lo && i < row_hi) { status = NSO_E_RANGE; goto out; }
                continue;
            }
            if (!t) continue;
            long long kv = ku + (long long)DI[d] * W + DJ[d];
            int vi = (int)(kv / W);
            double cand = g[ku] + (card ? 1 : M_SQRT2);
            card = !card;
            if (vi < row_lo || vi >= row_hi) continue;
            if (cand > cap) continue;              /* beyond the band: stays pending (see nso_field_min_pending) */
            if (g[kv] > cand + slack) {
                g[kv] = cand; improved++;
                if (!inq[kv]) { fifo[(head + cnt) % N] = (int32_t)kv; cnt++; inq[kv] = 1; }
            }
        }
    }
out:
    free(fifo); free(inq);
    return status < 0 ? status : improved;
}

/* Smallest candidate that would still improve a label of the shard (1e300 when it is at its fixed point): what a capped
 * nso_field_relax_window left pending. */
double nso_field_min_pending(const nso_map* m, const double* g, int row_lo, int row_hi, double slack) {
    const int H = m->H, W = m->W;
    double best = 1e300;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const long long ku = (long long)i * W + j;
            if (!(g[ku] < G_INF)) continue;
            int card = 1;
            for (int d = 0; d < 8; ++d) {
                int t = nso_test_pos(m, i + DI[d], j + DJ[d]);
                if (t == NSO_E_RANGE || !t) continue;
                long long kv = ku + (long long)DI[d] * W + DJ[d];
                int vi = (int)(kv / W);
                double cand = g[ku] + (card ? 1 : M_SQRT2);
                card = !card;
                if (vi < row_lo || vi >= row_hi) continue;
                if (g[kv] > cand + slack && cand < best) best = cand;
            }
        }
    return best;
}

/* ---- map producer, one step before the planner path -------------------------------------------
 * generateMask (occupancy_grid.cpp:73-89): the (i, j) offsets with i*i + j*j <= r^2 / cell^2, i and j looped as doubles from
 * -round(r / cell) to +round(r / cell).  Returns the number of offsets written to ab (pairs). */
int nso_occ_mask(double inflation_radius, double cell_size, int32_t* ab, int cap) {
    const double cells_in_radius = round(inflation_radius / cell_size);
    const double furthest_away = inflation_radius * inflation_radius / cell_size / cell_size;
    int n = 0;
    for (double i = -cells_in_radius; i <= cells_in_radius; ++i)
        for (double j = -cells_in_radius; j <= cells_in_radius; ++j)
            if (i * i + j * j <= furthest_away) {
                if (n < cap) { ab[2 * n] = (int32_t)i; ab[2 * n + 1] = (int32_t)j; }
                n++;
            }
    return n;
}

/* The inflation plane the reference holds for a given log-odds plane.  It keeps it incrementally (updateLogOdds /
 * updateInflation, occupancy_grid.cpp:166-213: +1 over the mask when a cell reaches thresh, -1 when it drops below), so at
 * any time infl[k] = number of cells at or above thresh whose mask reaches k, under the reference's index rules:
 * oob() tests only the row and with "<= 0" (220-223), so row 0 is never updated and never a target; the target key is
 * (ci + a) * W + (cj + b) with an unchecked column (flatten, 215-218); a key outside [0, N) makes vector::at throw.
 * Returns NSO_OK, or NSO_E_RANGE when some occupied cell's update would have thrown. */
int nso_occ_inflation(int H, int W, const int32_t* lo, int thresh, const int32_t* ab, int n_mask, int32_t* infl) {
    const long long N = (long long)H * W;
    int status = NSO_OK;
    for (long long k = 0; k < N; ++k) infl[k] = 0;
    for (int ci = 1; ci < H; ++ci)                       /* updateLogOdds returns at once for row 0 (oob) */
        for (int cj = 0; cj < W; ++cj) {
            if (lo[(long long)ci * W + cj] < thresh) continue;
            for (int m = 0; m < n_mask; ++m) {
                const int ti = ci + ab[2 * m], tj = cj + ab[2 * m + 1];
                if (ti <= 0 || ti >= H) continue;        /* oob(to_inflate): the column is never tested */
                const long long key = (long long)ti * W + tj;
                if (key < 0 || key >= N) { status = NSO_E_RANGE; continue; }
                infl[key] += 1;
            }
        }
    return status;
}


/* ---- the path consumer's safety check -----------------------------------------------------------
 * Commander::checkCell (commander.cpp:361-384): safe iff !oob && infl[k] <= 0 && lo[k] <= lo_thresh, with the commander's OWN
 * oob (394-397): `idx.i <= 0 || idx.i >= H || (idx.j < 0 && idx.j >= W)` - row 0 counts as outside, the column is never
 * tested, so the key i*W + j may alias another row or leave [0, N), where vector::at throws.
 * Returns 1 safe, 0 not, NSO_E_RANGE where the reference throws. */
int nso_cmd_check_cell(const nso_map* m, int i, int j) {
    if (i <= 0 || i >= m->H) return 0;
    const long long k = (long long)i * m->W + j;
    if (k < 0 || k >= (long long)m->H * m->W) return NSO_E_RANGE;
    if (m->infl[k] > 0) return 0;
    if (m->lo[k] > m->lo_thresh) return 0;
    return 1;
}

/* Commander::checkTrajectorySafety (commander.cpp:316-357) on trajectory cells ij (pos2idx of the trajectory points,
 * commander.cpp:386-391 == nso_pos2idx).  Empty trajectory: (unsafe, -2).  One point: that point, index -1 either way.
 * Otherwise points t_id, t_id-1, ..., 0 in that order; the first unsafe one is returned; trajectory_.at(i) throws for
 * i >= n (and t_id < 0 checks nothing).  Returns NSO_OK or NSO_E_RANGE. */
int nso_cmd_check_trajectory(const nso_map* m, int n, const int32_t* ij, int t_id, int* safe, int* bad_idx) {
    if (n == 0) { *safe = 0; *bad_idx = -2; return NSO_OK; }
    if (n == 1) {
        const int s = nso_cmd_check_cell(m, ij[0], ij[1]);
        if (s < 0) return s;
        *safe = s; *bad_idx = -1;
        return NSO_OK;
    }
    for (int i = t_id; i >= 0; --i) {
        if (i >= n) return NSO_E_RANGE;                 /* trajectory_.at(i) */
        const int s = nso_cmd_check_cell(m, ij[2 * i], ij[2 * i + 1]);
        if (s < 0) return s;
        if (!s) { *safe = 0; *bad_idx = i; return NSO_OK; }
    }
    *safe = 1; *bad_idx = -1;
    return NSO_OK;
}

/* ---- the opt-in textbook graph ------------------------------------------------------------------
 * NOT reference behaviour (there is none to match): the cost-to-go field a textbook Dijkstra computes on the plain
 * 8-connected grid over the same free cells: neighbours bounds-checked in rows AND columns, cardinal steps 1, diagonal steps
 * sqrt(2), no corner rule (like the reference, a diagonal may pass between two blocked cells).  fp64, binary heap with lazy
 * deletion; g = 1e5 where unreached, as the reference initialises it.  Checks the CUDA field on the "graph = 1" masks. */
typedef struct { double g; long long k; } tb_ent;
static void tb_push(tb_ent* h, long long* n, double g, long long k) {
    long long c = (*n)++;
    while (c > 0) {
        long long p = (c - 1) / 2;
        if (h[p].g <= g) break;
        h[c] = h[p]; c = p;
    }
    h[c].g = g; h[c].k = k;
}
static tb_ent tb_pop(tb_ent* h, long long* n) {
    tb_ent top = h[0], last = h[--(*n)];
    long long c = 0;
    for (;;) {
        long long l = 2 * c + 1;
        if (l >= *n) break;
        if (l + 1 < *n && h[l + 1].g < h[l].g) l++;
        if (h[l].g >= last.g) break;
        h[c] = h[l]; c = l;
    }
    if (*n > 0) h[c] = last;
    return top;
}
int nso_textbook_field(const nso_map* m, int si, int sj, double* g) {
    static const int DI[8] = {1, 1, 0, -1, -1, -1, 0, 1}, DJ[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    const int H = m->H, W = m->W;
    const long long N = (long long)H * W;
    if (si < 0 || si >= H || sj < 0 || sj >= W) return NSO_E_ARG;
    tb_ent* heap = (tb_ent*)malloc(sizeof(tb_ent) * (size_t)(8 * N + 8));
    if (!heap) return NSO_E_NOMEM;
    long long n = 0;
    for (long long k = 0; k < N; ++k) g[k] = 1e5;
    g[(long long)si * W + sj] = 0.0;
    tb_push(heap, &n, 0.0, (long long)si * W + sj);
    while (n > 0) {
        const tb_ent e = tb_pop(heap, &n);
        if (e.g > g[e.k]) continue;
        const int i = (int)(e.k / W), j = (int)(e.k % W);
        for (int d = 0; d < 8; ++d) {
            const int ni = i + DI[d], nj = j + DJ[d];
            if (ni < 0 || ni >= H || nj < 0 || nj >= W) continue;
            const long long kv = (long long)ni * W + nj;
            if (m->infl[kv] > 0 || m->lo[kv] > m->lo_thresh) continue;
            const double cand = e.g + ((d & 1) ? M_SQRT2 : 1.0);
            if (cand < g[kv]) { g[kv] = cand; tb_push(heap, &n, cand, kv); }
        }
    }
    free(heap);
    return NSO_OK;
}
