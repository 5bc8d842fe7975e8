/* TEST INFRASTRUCTURE ONLY — CPU restatement of the reference planner algorithms (see nsfs_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it. */
#ifndef NSFS_ORACLE_H
#define NSFS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int H, W;               /* map_size_.i, map_size_.j  (map_data.h:17) */
    const int32_t* infl;    /* grid_inflation_, row-major key = i*W + j  (map_data.h:10) */
    const int32_t* lo;      /* grid_logodds_   (map_data.h:11) */
    int lo_thresh;          /* lo_thresh_      (map_data.h:12) */
    double cell, ox, oy;    /* cell_size_, origin_ (map_data.h:14,18) */
} nso_map;

typedef struct {
    long long pops;         /* OpenList::popNode calls */
    long long expansions;   /* non-stale pops == cells marked visited ("settled") */
    long long pushes;       /* OpenList::pushNode calls */
    long long heap_peak;    /* largest queue length */
    long long los_checks;   /* isLos calls (ThetaStar only) */
} nso_stats;

enum { NSO_ASTAR = 0, NSO_DJIKSTRA = 1, NSO_THETASTAR = 2 };
enum { NSO_MODE_F = 0, NSO_MODE_G = 1, NSO_MODE_FG = 2 };
enum { NSO_OK = 0, NSO_E_RANGE = -1, NSO_E_TRUNC = -3, NSO_E_ARG = -4, NSO_E_NOMEM = -5 };

/* plan(): raw path as (i,j) pairs ordered goal -> start; unreachable => the single start cell.
 * g_out/parent_out/visited_out (each H*W, optional) receive the node store after the search. */
int nso_plan(const nso_map* m, int algo, int mode, int si, int sj, int gi, int gj,
             int32_t* path_ij, int cap, int* n_path, double* g_goal,
             double* g_out, int32_t* parent_out, uint8_t* visited_out, nso_stats* st);

int nso_find_better_point(const nso_map* m, int bi, int bj, int* fi, int* fj, nso_stats* st);

int nso_test_pos(const nso_map* m, int i, int j);               /* 1 free, 0 not, NSO_E_RANGE */
int nso_bresenham(int si, int sj, int ti, int tj, int32_t* out_ij, int cap);    /* returns #cells */
int nso_bresenham_closed(int si, int sj, int ti, int tj, int32_t* out_ij, int cap);
int nso_is_los(const nso_map* m, int si, int sj, int ti, int tj); /* 1, 0, NSO_E_RANGE */

double nso_dist_oct(int si, int sj, int ti, int tj);
double nso_dist_euc(int si, int sj, int ti, int tj);
void nso_idx2pos(const nso_map* m, int i, int j, double* x, double* y);
void nso_pos2idx(const nso_map* m, double x, double y, int* i, int* j);

int nso_sparsify(const double* xy_in, int n_in, double* xy_out, int cap);
int nso_any_angle(const nso_map* m, const double* xy_in, int n_in, double* xy_out, int cap);
int nso_post_process(const nso_map* m, const double* xy_in, int n_in, double* xy_out, int cap);

/* OpenList trace: ops[4*t..] = {op, f, g, id}; op 1 push, op 0 pop (popped id appended to out). */
int nso_heap_trace(int mode, int n_ops, const int32_t* ops, int32_t* out);

/* Order-independent check: label-correcting fixed point of the plan() edge rule. */
int nso_field_fixed_point(const nso_map* m, int si, int sj, double* g);

/* One shard of a row-partitioned field: warm start from g, rows outside [row_lo,row_hi) take no in-edges. */
long long nso_field_relax_window(const nso_map* m, double* g, int row_lo, int row_hi, double slack, double cap);
double nso_field_min_pending(const nso_map* m, const double* g, int row_lo, int row_hi, double slack);

/* Map producer (reference loco_mapping): disc mask and the inflation plane implied by a log-odds plane. */
int nso_occ_mask(double inflation_radius, double cell_size, int32_t* ab, int cap);
/* opt-in textbook graph (not reference behaviour): plain 8-connected Dijkstra field, fp64, 1e5 = unreached */
int nso_textbook_field(const nso_map* m, int si, int sj, double* g);

/* the path consumer's safety check (commander.cpp:316-402) */
int nso_cmd_check_cell(const nso_map* m, int i, int j);           /* 1 safe, 0 not, NSO_E_RANGE */
int nso_cmd_check_trajectory(const nso_map* m, int n, const int32_t* ij, int t_id, int* safe, int* bad_idx);

int nso_occ_inflation(int H, int W, const int32_t* lo, int thresh, const int32_t* ab, int n_mask, int32_t* infl);

#ifdef __cplusplus
}
#endif
#endif
