/*
 * rbpf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, float64, CPU restatement of the per-scan RBPF update of
 * amansanghvi/Thesis.  It exists to check the CUDA path (thesis_b200/csrc) and
 * to serve as the timed CPU baseline.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; nothing under
 * thesis_b200/ does.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the reference checkout).  Data layout deliberately follows the reference, not
 * the GPU: a map is a list of 40 m x 40 m tiles of 800x800 float64 log-odds in
 * allocation order (hybridmap.py:63-70, gridmap.py:31-32), a duplicate particle
 * is a deep copy (robot.py:141-149).
 *
 * Parity status
 *   - pinned against the reference's own Python (imported with stubs, see
 *     oracle/ref_shim.py and tests/golden/make_golden.py): beam geometry, tile
 *     lookup, ray-cast map integration, sample weights, moments, resampling,
 *     motion models, nearby-occupied gather.
 *   - PARITY UNPINNED: orc_match().  The reference delegates the search to
 *     MathWorks matchScansGrid/matchScans (matchScanCustom.m:9-37), whose source
 *     is not in the reference tree and cannot run here.  orc_match() is OUR
 *     restatement of a correlative grid search honouring the call contract
 *     (matchScanCustom.m:1-58, hybridmap.py:210-261); see DESIGN.md section 4.
 *
 * Declared deviations from the reference (same ones the GPU path makes):
 *   - np.longdouble accumulators (robot.py:25,92-94,119,124) are float64.
 *   - every particle owns a private tile list (the class-level list at
 *     hybridmap.py:64 aliases all particles' maps).
 *   - proposal samples use the lower Cholesky factor of the matcher covariance
 *     on caller-supplied standard normals instead of NumPy's SVD transform
 *     (robot.py:81); same distribution, different individual draws.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* sin / cos / exp in plain IEEE operations, shared with the CUDA kernels so that weights (and with
 * them the resampled ancestors) agree bit for bit; within 1 ulp of libm, equal to it for > 99.8 % of
 * arguments (tests/test_rb_math.py). */
#include "../thesis_b200/csrc/rb_math.h"

/* the shared routines, exported for tests/test_rb_math.py */
void orc_rb_sincos(const double *a, int n, double *sn, double *cs) { for (int i = 0; i < n; i++) rb_sincos(a[i], &sn[i], &cs[i]); }
void orc_rb_exp(const double *x, int n, double *out) { for (int i = 0; i < n; i++) out[i] = rb_exp(x[i]); }

#define CS 0.05            /* hybridmap.py:67  cell size, metres */
#define TILE_LEN 40        /* hybridmap.py:68  tile side, metres (an int in the reference) */
#define DIM 800            /* gridmap.py:31    round(40/0.05) */
#define L_OCC 0.80         /* gridmap.py:20 */
#define L_NEAR 0.20        /* gridmap.py:21 */
#define L_MAX 3.0          /* gridmap.py:22 */
#define L_EMP (-0.30)      /* gridmap.py:23 */
#define L_MIN (-3.0)       /* gridmap.py:24 */
#define MATCH_MAX_R 11.0   /* hybridmap.py:20  VALID_DIST_THRESHOLD */
#define MATCH_MIN_R 1e-3   /* hybridmap.py:218 */
#define CLIP_R 15.0        /* hybridmap.py:107-108 */
#define W_MAX_R 25.0       /* robot.py:130 */
#define W_MIN_R 0.01       /* robot.py:130 */
#define RESAMPLE_TRIGGER 200.0 /* main.py:50 */
#define ROT_RANGE (M_PI / 6.0) /* hybridmap.py:249 */
#define MAX_NT 14          /* window clamp 0.7 m (robot.py:64-65) / 0.05 m */
#define ROT_LINE_HALF 16   /* rotation-variance support: +-16 lattice steps (~4.2 deg) around the best rotation */

/* ------------------------------------------------------------------ map -- */

typedef struct {
    int cx, cy;            /* tile centre, metres, on the 40 m lattice (hybridmap.py:193-208) */
    double *cells;         /* [ix*DIM + iy], x is the row index (gridmap.py:87) */
} orc_tile;

typedef struct {
    orc_tile *tiles;       /* allocation order == lookup order (hybridmap.py:269-272) */
    int n, cap;
} orc_map;

static void map_push(orc_map *m, int cx, int cy)
{
    if (m->n == m->cap) {
        m->cap = m->cap ? 2 * m->cap : 4;
        m->tiles = (orc_tile *)realloc(m->tiles, (size_t)m->cap * sizeof(orc_tile));
    }
    m->tiles[m->n].cx = cx;
    m->tiles[m->n].cy = cy;
    m->tiles[m->n].cells = (double *)calloc((size_t)DIM * DIM, sizeof(double));
    m->n++;
}

/* HybridMap.__init__ hybridmap.py:66-70 : one blank tile centred on (0,0). */
orc_map *orc_map_new(void)
{
    orc_map *m = (orc_map *)calloc(1, sizeof(orc_map));
    map_push(m, 0, 0);
    return m;
}

void orc_map_free(orc_map *m)
{
    if (!m) return;
    for (int i = 0; i < m->n; i++) free(m->tiles[i].cells);
    free(m->tiles);
    free(m);
}

/* HybridMap.copy hybridmap.py:315-320 -> HybridMapEntry.copy :56-61 -> GridMap.copy
 * gridmap.py:336-342 : deep copy of every tile. */
orc_map *orc_map_copy(const orc_map *src)
{
    orc_map *m = (orc_map *)calloc(1, sizeof(orc_map));
    m->cap = src->n > 4 ? src->n : 4;
    m->tiles = (orc_tile *)malloc((size_t)m->cap * sizeof(orc_tile));
    for (int i = 0; i < src->n; i++) {
        m->tiles[i].cx = src->tiles[i].cx;
        m->tiles[i].cy = src->tiles[i].cy;
        m->tiles[i].cells = (double *)malloc((size_t)DIM * DIM * sizeof(double));
        memcpy(m->tiles[i].cells, src->tiles[i].cells, (size_t)DIM * DIM * sizeof(double));
    }
    m->n = src->n;
    return m;
}

int orc_map_ntiles(const orc_map *m) { return m->n; }
void orc_map_tile_centre(const orc_map *m, int i, int *cx, int *cy)
{
    *cx = m->tiles[i].cx;
    *cy = m->tiles[i].cy;
}
double *orc_map_tile_cells(orc_map *m, int i) { return m->tiles[i].cells; }

/* Tile with this centre, or NULL. */
double *orc_map_find_tile(orc_map *m, int cx, int cy)
{
    for (int i = 0; i < m->n; i++)
        if (m->tiles[i].cx == cx && m->tiles[i].cy == cy) return m->tiles[i].cells;
    return NULL;
}

/* HybridMapEntry.is_in_map hybridmap.py:44-45 with the bounds of :31-36. */
static int tile_has(const orc_tile *t, double x, double y)
{
    double r = TILE_LEN / 2.0;
    return x >= t->cx - r && x < t->cx + r && y >= t->cy - r && y < t->cy + r;
}

/* HybridMap.get_map_with_pos hybridmap.py:263-272 : first match in list order. */
static orc_tile *map_with_pos(const orc_map *m, double x, double y)
{
    for (int i = 0; i < m->n; i++)
        if (tile_has(&m->tiles[i], x, y)) return &m->tiles[i];
    return NULL;
}

/* GridMap.get_cell gridmap.py:120-128 (tile-relative metres -> indices). */
static int get_cell(double x, double y, int *ix, int *iy)
{
    double h = TILE_LEN / 2.0;
    if (y < -h || y >= h) return 0;
    if (x < -h || x >= h) return 0;
    *ix = (int)(x / TILE_LEN * DIM + DIM / 2.0);
    *iy = (int)(y / TILE_LEN * DIM + DIM / 2.0);
    return 1;
}

/* GridMap.index_to_distance gridmap.py:333-334. */
static double index_to_distance(int i) { return (double)(i - DIM / 2.0) * TILE_LEN / DIM; }

/* HybridMap.get_odds_at hybridmap.py:85-93.  Returns 1 and *out when a tile
 * holds the point, 0 for the reference's None. */
int orc_get_odds_at(const orc_map *m, double x, double y, double *out)
{
    orc_tile *t = map_with_pos(m, x, y);
    int ix, iy;
    if (!t) return 0;
    if (!get_cell(x - t->cx, y - t->cy, &ix, &iy)) return 0;
    *out = t->cells[(size_t)ix * DIM + iy];
    return 1;
}

/* HybridMap._get_map_centre hybridmap.py:193-208. */
static void map_centre(double x, double y, int *cx, int *cy)
{
    int ax = (int)nearbyint(x / TILE_LEN);     /* Python round(): half to even, like nearbyint */
    int ay = (int)nearbyint(y / TILE_LEN);
    *cx = 0;
    *cy = 0;
    for (int a = ax - 1; a < ax + 2; a++) {
        int mc = a * TILE_LEN;
        if (x < mc + TILE_LEN / 2.0 && x >= mc - TILE_LEN / 2.0) { *cx = mc; break; }
    }
    for (int a = ay - 1; a < ay + 2; a++) {
        int mc = a * TILE_LEN;
        if (y < mc + TILE_LEN / 2.0 && y >= mc - TILE_LEN / 2.0) { *cy = mc; break; }
    }
}

/* Python negative indices wrap (ndarray[-1]); keep that behaviour visible. */
static size_t wrap_idx(int ix, int iy)
{
    if (ix < 0) ix += DIM;
    if (iy < 0) iy += DIM;
    return (size_t)ix * DIM + iy;
}

/* GridMap.set_occupied_pos / set_empty_pos / set_nearby_pos gridmap.py:86-117. */
static void set_pos(orc_tile *t, double rx, double ry, int kind)
{
    int ix = (int)(rx / CS + DIM / 2.0);
    int iy = (int)(ry / CS + DIM / 2.0);
    double *c = &t->cells[wrap_idx(ix, iy)];
    if (kind == 0) *c = fmax(*c + L_EMP, L_MIN);
    else if (kind == 1) *c = fmin(*c + L_OCC, L_MAX);
    else *c = fmin(*c + L_NEAR, L_MAX);
}

/* ---------------------------------------------------------------- beams -- */

/* Scan.__init__ lidar.py:76-80 plus the per-beam range used by the gates
 * (hybridmap.py:105,217; robot.py:129). */
void orc_scan_prepare(const double *ranges, const double *angles, int B,
                      double *px, double *py, double *dist)
{
    for (int j = 0; j < B; j++) {
        px[j] = ranges[j] * cos(angles[j]);
        py[j] = ranges[j] * sin(angles[j]);
        dist[j] = sqrt(px[j] * px[j] + py[j] * py[j]);
    }
}

/* Scan.from_global_reference lidar.py:111-128 : [c -s x; s c y; 0 0 1] . [px;py;1]. */
static inline void xform(double c, double s, double x, double y, double px, double py,
                         double *gx, double *gy)
{
    /* np.matmul of the 3x3 by 3xB (lidar.py:123) accumulates k = 0,1,2 with fused
     * multiply-adds on the build container's BLAS: round(c*px), fma(-s,py,.), + x.
     * Replaying that order makes endpoints bit-identical to the reference there. */
    *gx = fma(-s, py, c * px) + x;
    *gy = fma(c, py, s * px) + y;
}

void orc_transform(const double *pose, const double *px, const double *py, int B,
                   double *gx, double *gy)
{
    double c, s;
    rb_sincos(pose[2], &s, &c);
    for (int j = 0; j < B; j++) xform(c, s, pose[0], pose[1], px[j], py[j], &gx[j], &gy[j]);
}

/* ------------------------------------------------------ map integration -- */

/* HybridMap.get_affected_points hybridmap.py:274-301, including the empty list
 * for axis-aligned rays that point in the negative direction (:278-281).
 * Returns the number of cells written to out (capacity cap). */
int orc_bresenham(int x0, int y0, int x1, int y1, int *out, int cap)
{
    int dx = abs(x1 - x0), dy = abs(y1 - y0), n = 0;
    if (dx == 0) {
        for (int y = y0; y < y1 + 1; y++) { if (n < cap) { out[2 * n] = x0; out[2 * n + 1] = y; } n++; }
        return n;
    }
    if (dy == 0) {
        for (int x = x0; x < x1 + 1; x++) { if (n < cap) { out[2 * n] = x; out[2 * n + 1] = y0; } n++; }
        return n;
    }
    int xs = x1 - x0 > 0 ? 1 : -1, ys = y1 - y0 > 0 ? 1 : -1;
    int steep = dy > dx;
    if (steep) { int t = dx; dx = dy; dy = t; }
    int D = 2 * dy - dx, y = 0;
    for (int x = 0; x < dx + 1; x++) {
        if (n < cap) {
            if (steep) { out[2 * n] = x0 + xs * y; out[2 * n + 1] = y0 + ys * x; }
            else       { out[2 * n] = x0 + xs * x; out[2 * n + 1] = y0 + ys * y; }
        }
        n++;
        if (D >= 0) { y += 1; D -= 2 * dx; }
        D += 2 * dy;
    }
    return n;
}

/* HybridMap.update hybridmap.py:95-145. */
void orc_map_update(orc_map *m, const double *pose, const double *px, const double *py,
                    const double *dist, int B)
{
    if (!map_with_pos(m, pose[0], pose[1])) return;               /* :98-100 */
    int sx = (int)(pose[0] / CS), sy = (int)(pose[1] / CS);      /* :102 */
    double c, s;
    rb_sincos(pose[2], &s, &c);
    int cap = 4096, *pts = (int *)malloc(sizeof(int) * 2 * (size_t)cap);
    for (int j = 0; j < B; j++) {
        double gx, gy;
        xform(c, s, pose[0], pose[1], px[j], py[j], &gx, &gy);
        int occ = 1;
        int ex = (int)(gx / CS), ey = (int)(gy / CS);             /* :106 */
        if (dist[j] > CLIP_R) {                                   /* :107-113 */
            double scale = 15.0 / dist[j];
            int nex = (int)(sx + scale * (ex - sx));
            int ney = (int)(sy + scale * (ey - sy));
            ex = nex; ey = ney; occ = 0;
        }
        int n = orc_bresenham(sx, sy, ex, ey, pts, cap);
        if (n > cap) {
            cap = n; pts = (int *)realloc(pts, sizeof(int) * 2 * (size_t)cap);
            n = orc_bresenham(sx, sy, ex, ey, pts, cap);
        }
        for (int q = 0; q < n; q++) {                             /* :122-144 */
            double X = pts[2 * q] * CS, Y = pts[2 * q + 1] * CS;
            orc_tile *t = map_with_pos(m, X, Y);
            if (!t) {
                int cx, cy;
                map_centre(X, Y, &cx, &cy);
                t = map_with_pos(m, (double)cx, (double)cy);
                if (!t) { map_push(m, cx, cy); t = &m->tiles[m->n - 1]; }
            }
            if (occ && pts[2 * q] == ex && pts[2 * q + 1] == ey) {
                set_pos(t, X - t->cx, Y - t->cy, 1);
                if (q > 0) {
                    double NX = pts[2 * q - 2] * CS, NY = pts[2 * q - 1] * CS;
                    if (tile_has(t, NX, NY)) set_pos(t, NX - t->cx, NY - t->cy, 2);
                }
            } else {
                set_pos(t, X - t->cx, Y - t->cy, 0);
            }
        }
    }
    free(pts);
}

/* GridMap.get_nearby_occ_points via HybridMap.get_scan_match hybridmap.py:230-234
 * and gridmap.py:130-155 : cells with log-odds > 1.0 in the 72x72 window around
 * a point, over every tile.  Writes (x,y) metre pairs; returns the count. */
int orc_nearby_occ(const orc_map *m, double cpx, double cpy, double *out, int cap)
{
    int n = 0, pr = (int)(1.8 / CS);
    for (int i = 0; i < m->n; i++) {
        const orc_tile *t = &m->tiles[i];
        double x = cpx - t->cx, y = cpy - t->cy;
        int decx = 0, decy = 0;
        if (y < -TILE_LEN / 2.0) decy = 1; else if (x < -TILE_LEN / 2.0) decx = 1;   /* :131-136 */
        int jx = (int)(x / TILE_LEN * DIM + DIM / 2.0) - decx;
        int jy = (int)(y / TILE_LEN * DIM + DIM / 2.0) - decy;
        int x0 = jx - pr > 0 ? jx - pr : 0, y0 = jy - pr > 0 ? jy - pr : 0;
        int x1 = jx + pr < DIM ? jx + pr : DIM, y1 = jy + pr < DIM ? jy + pr : DIM;
        for (int a = x0; a < x1; a++)
            for (int b = y0; b < y1; b++)
                if (t->cells[(size_t)a * DIM + b] > 1.0) {
                    if (n < cap) { out[2 * n] = index_to_distance(a) + t->cx; out[2 * n + 1] = index_to_distance(b) + t->cy; }
                    n++;
                }
    }
    return n;
}

/* ------------------------------------------------------------- weights -- */

/* Robot._generate_sample_weight robot.py:118-139 for K guesses. */
void orc_sample_weight(const orc_map *m, const double *guesses, int K, const double *px,
                       const double *py, const double *dist, int B, const double *prs, double *w)
{
    for (int k = 0; k < K; k++) {
        const double *g = guesses + 3 * k;
        double c, s;
        long tenths = 10;                                                   /* the 1 + ... of robot.py:135 */
        rb_sincos(g[2], &s, &c);
        for (int j = 0; j < B; j++) {
            if (dist[j] < W_MAX_R && dist[j] > W_MIN_R) {
                double gx, gy, L;
                xform(c, s, g[0], g[1], px[j], py[j], &gx, &gy);
                /* Declared deviation: the log-odds are summed exactly, in tenths (a cell is a multiple of
                 * 0.1 to within 1.1e-14), where the reference adds the float64 cells; the
                 * two sums differ by < B * 1.1e-14.  The CUDA path holds int8 tenths and forms the same
                 * integer, which makes the weights -- and the resampled ancestors -- agree bit for bit. */
                if (orc_get_odds_at(m, gx, gy, &L)) tenths += lrint(L * 10.0);
            }
        }
        w[k] = ((double)tenths / 10.0) * prs[k];
    }
}

/* Lower Cholesky factor of a symmetric 3x3 (row-major); our sampling transform. */
static void chol3(const double *c, double *l)
{
    l[0] = sqrt(c[0]);
    l[1] = c[3] / l[0];
    l[2] = c[6] / l[0];
    l[3] = sqrt(c[4] - l[1] * l[1]);
    l[4] = (c[7] - l[2] * l[1]) / l[3];
    l[5] = sqrt((c[8] - l[2] * l[2]) - l[4] * l[4]);
}

/* guesses = mean + L z  (stands in for np.random.multivariate_normal, robot.py:81)
 * prs     = N(guess; mean, cov) * 10   (robot.py:87) */
void orc_propose(const double *mean, const double *cov, const double *z, int K,
                 double *guesses, double *prs)
{
    double l[6];
    chol3(cov, l);
    double nrm = (2.0 * M_PI) * sqrt(2.0 * M_PI) * ((l[0] * l[3]) * l[5]);
    for (int k = 0; k < K; k++) {
        const double *zz = z + 3 * k;
        double *g = guesses + 3 * k;
        g[0] = mean[0] + l[0] * zz[0];
        g[1] = mean[1] + (l[1] * zz[0] + l[3] * zz[1]);
        g[2] = mean[2] + ((l[2] * zz[0] + l[4] * zz[1]) + l[5] * zz[2]);
        double d0 = g[0] - mean[0], d1 = g[1] - mean[1], d2 = g[2] - mean[2];
        double y0 = d0 / l[0];
        double y1 = (d1 - l[1] * y0) / l[3];
        double y2 = ((d2 - l[2] * y0) - l[4] * y1) / l[5];
        double maha = (y0 * y0 + y1 * y1) + y2 * y2;
        prs[k] = rb_exp(-0.5 * maha) / nrm * 10.0;
    }
}

/* prs = N(guess; mean, cov) * 10 (robot.py:87) for samples the caller drew itself -- with
 * np.random.multivariate_normal like the reference (robot.py:81). */
void orc_propose_pdf(const double *mean, const double *cov, const double *guesses, int K, double *prs)
{
    double l[6];
    chol3(cov, l);
    double nrm = (2.0 * M_PI) * sqrt(2.0 * M_PI) * ((l[0] * l[3]) * l[5]);
    for (int k = 0; k < K; k++) {
        const double *g = guesses + 3 * k;
        double d0 = g[0] - mean[0], d1 = g[1] - mean[1], d2 = g[2] - mean[2];
        double y0 = d0 / l[0];
        double y1 = (d1 - l[1] * y0) / l[3];
        double y2 = ((d2 - l[2] * y0) - l[4] * y1) / l[5];
        double maha = (y0 * y0 + y1 * y1) + y2 * y2;
        prs[k] = rb_exp(-0.5 * maha) / nrm * 10.0;
    }
}

/* Weight normalisation and weighted moments, robot.py:88-108.  Returns norm. */
double orc_moments(const double *guesses, const double *w, int K, double *mean, double *sigma)
{
    double mn = w[0];
    for (int k = 1; k < K; k++) if (w[k] < mn) mn = w[k];
    double norm = 0.0;
    mean[0] = mean[1] = mean[2] = 0.0;
    for (int k = 0; k < K; k++) {
        double wk = (w[k] - mn) + 1e-2;
        for (int a = 0; a < 3; a++) mean[a] = mean[a] + guesses[3 * k + a] * wk;
        norm = norm + wk;
    }
    for (int a = 0; a < 3; a++) mean[a] = mean[a] / norm;
    for (int a = 0; a < 9; a++) sigma[a] = 0.0;
    for (int k = 0; k < K; k++) {
        double wk = (w[k] - mn) + 1e-2, d[3];
        for (int a = 0; a < 3; a++) d[a] = guesses[3 * k + a] + (-mean[a]);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) sigma[3 * a + b] = sigma[3 * a + b] + (d[a] * d[b]) * wk;
    }
    for (int a = 0; a < 9; a++) sigma[a] = sigma[a] / norm;
    return norm + mn * K;
}

/* ------------------------------------------------------------- matcher -- */

/* Rotation lattice of OUR restated matcher: step = acos(1 - d^2/(2 r^2)) with
 * d = cell size and r = the matcher's range gate, so that a point at the gate
 * moves at most one cell per step. */
double orc_rot_step(void) { return acos(1.0 - (CS * CS) / (2.0 * MATCH_MAX_R * MATCH_MAX_R)); }
int orc_rot_count(void) { return (int)floor(ROT_RANGE / orc_rot_step()); }

/* Window half-width in cells for a clamp value from robot.py:64-65. */
int orc_window_cells(double r)
{
    int n = (int)floor(r / CS + 1e-9);
    return n > MAX_NT ? MAX_NT : n;
}

/* Robot.map_update robot.py:62-65 : search window from the pose covariance. */
void orc_pose_range(const double *cov, double *rx, double *ry)
{
    double p0 = sqrt(cov[0]) * 30.0, p1 = sqrt(cov[4]) * 30.0;
    *rx = fmax(fmin(4 * p0, 0.7), 0.1);
    *ry = fmax(fmin(4 * p1, 0.7), 0.1);
}

/* Integer "tenths" of a log-odds value (values are sums of +0.8/+0.2/-0.3). */
static int tenths(double L) { return (int)lround(L * 10.0); }

/* Occupancy of a cell of the global read lattice: G = 800*t + ix - 400. */
static int occ_cell(const orc_map *m, int Gx, int Gy)
{
    int tx = (int)floor((Gx + 400) / 800.0), ty = (int)floor((Gy + 400) / 800.0);
    int ix = Gx + 400 - 800 * tx, iy = Gy + 400 - 800 * ty;
    for (int i = 0; i < m->n; i++)
        if (m->tiles[i].cx == tx * TILE_LEN && m->tiles[i].cy == ty * TILE_LEN)
            return tenths(m->tiles[i].cells[(size_t)ix * DIM + iy]) > 10;    /* gridmap.py:153 (L > 1.0) */
    return 0;
}

static int64_t match_key(int score, int i, int j, int k)
{
    /* higher score first; then smaller |k|, |i|, |j|; then negative before positive */
    int64_t key = (int64_t)score << 32;
    key |= (int64_t)(255 - abs(k)) << 24;
    key |= (int64_t)(31 - abs(i)) << 19;
    key |= (int64_t)(31 - abs(j)) << 14;
    key |= (int64_t)(k < 0) << 13;
    key |= (int64_t)(i < 0) << 12;
    key |= (int64_t)(j < 0) << 11;
    return key;
}

/*
 * Restated scan-to-map matcher (G1 + G2 of SURVEY section 8a).
 *   front-end  hybridmap.py:210-240 : curr points = beam endpoints at the guess
 *              (range gate 1e-3 < r < 11), snapped to the corner of their cell in
 *              an existing tile, relative to the guess position, |c| < 11.
 *   search     contract of matchScanCustom.m:9-17 : correlative search over
 *              |dx|<=rx, |dy|<=ry on the 0.05 m lattice and |dth|<=pi/6 on the
 *              orc_rot_step() lattice; score = number of curr points whose
 *              rotated+shifted cell is occupied (tenths > 10) in the particle's
 *              own tiles.
 *   gate       matchScanCustom.m:52-57 isValidPose : strictly inside the window
 *              and not the zero correction, else cov = NaN, score = 0 (:26-28).
 *   covariance weights 2^(score - best) over the translation slice at the best
 *              rotation and over the rotation line (+-16 steps) at the best translation,
 *              cross terms zero, plus the lattice quantisation variance.
 *   result     guess + correction, hybridmap.py:253-255.
 *   refine     matchScanCustom.m:32-50 : the NDT stage above, when enabled with
 *              orc_set_refine(1) (off by default).
 * out_pose[3], out_cov[9]; returns 1 if valid else 0.  dbg (nullable) receives
 * {M, best_i, best_j, best_k, nx, ny, ndt_evals, ndt_accepted}; slice (nullable, 29*29 ints) the scores of
 * the translation slice at the best rotation, row j, column i.
 */
/* Front-end of HybridMap.get_scan_match hybridmap.py:216-228,236,240: the
 * `valid_curr_points` list handed to the matcher.  cx, cy have room for B. */
static int match_curr(const orc_map *m, const double *guess, const double *px, const double *py,
                      const double *dist, int B, double *cx, double *cy, int *cj)
{
    int M = 0;
    double c0, s0;
    rb_sincos(guess[2], &s0, &c0);
    for (int j = 0; j < B; j++) {
        if (!(dist[j] < MATCH_MAX_R && dist[j] > MATCH_MIN_R)) continue;      /* :217-218 */
        double gx, gy;
        xform(c0, s0, guess[0], guess[1], px[j], py[j], &gx, &gy);
        orc_tile *t = map_with_pos(m, gx, gy);                                /* :220-221 */
        int ix, iy;
        if (!t || !get_cell(gx - t->cx, gy - t->cy, &ix, &iy)) continue;
        double qx = (index_to_distance(ix) + t->cx) - guess[0];               /* :226-228,236 */
        double qy = (index_to_distance(iy) + t->cy) - guess[1];
        if (!(sqrt(qx * qx + qy * qy) < MATCH_MAX_R)) continue;               /* :240 */
        cx[M] = qx; cy[M] = qy; cj[M] = j; M++;
    }
    return M;
}

int orc_match_curr(const orc_map *m, const double *guess, const double *px, const double *py,
                   const double *dist, int B, double *out_xy)
{
    double *cx = (double *)malloc(sizeof(double) * (size_t)B), *cy = (double *)malloc(sizeof(double) * (size_t)B);
    int *cj = (int *)malloc(sizeof(int) * (size_t)B);
    int M = match_curr(m, guess, px, py, dist, B, cx, cy, cj);
    for (int q = 0; q < M; q++) { out_xy[2 * q] = cx[q]; out_xy[2 * q + 1] = cy[q]; }
    free(cx); free(cy); free(cj);
    return M;
}

static void match_ref_mask(const orc_map *m, const double *guess, const double *px, const double *py, const double *dist,
                           int B, int G0x, int G0y, int R1, unsigned char *mask);

/* The `valid_ref_points` of hybridmap.py:230-239 as the restated matcher sees them: occupied cells under
 * match_ref_mask, as (x, y) relative to the guess, in the lexicographic order of np.unique.  Returns the
 * count (out_xy may be NULL). */
int orc_match_ref(const orc_map *m, const double *guess, const double *px, const double *py, const double *dist, int B,
                  double *out_xy, int cap)
{
    int t0x = (int)floor(guess[0] / TILE_LEN + 0.5), t0y = (int)floor(guess[1] / TILE_LEN + 0.5);
    if (guess[0] < t0x * TILE_LEN - 20.0) t0x--; else if (guess[0] >= t0x * TILE_LEN + 20.0) t0x++;
    if (guess[1] < t0y * TILE_LEN - 20.0) t0y--; else if (guess[1] >= t0y * TILE_LEN + 20.0) t0y++;
    const int i0x = (int)((guess[0] - t0x * TILE_LEN) / TILE_LEN * DIM + DIM / 2.0);
    const int i0y = (int)((guess[1] - t0y * TILE_LEN) / TILE_LEN * DIM + DIM / 2.0);
    const int G0x = 800 * t0x + i0x - 400, G0y = 800 * t0y + i0y - 400, R1 = 300, S1 = 2 * R1 + 1;
    unsigned char *mask = (unsigned char *)malloc((size_t)S1 * S1);
    match_ref_mask(m, guess, px, py, dist, B, G0x, G0y, R1, mask);
    int n = 0;
    for (int a = -R1; a <= R1; a++)                                            /* x-major = lexicographic */
        for (int b = -R1; b <= R1; b++) {
            if (!mask[(size_t)(b + R1) * S1 + (a + R1)] || !occ_cell(m, G0x + a, G0y + b)) continue;
            const int Gx = G0x + a, Gy = G0y + b;
            const int tx = (int)floor((Gx + 400) / 800.0), ty = (int)floor((Gy + 400) / 800.0);
            if (out_xy && n < cap) {
                out_xy[2 * n] = (index_to_distance(Gx + 400 - 800 * tx) + tx * TILE_LEN) - guess[0];
                out_xy[2 * n + 1] = (index_to_distance(Gy + 400 - 800 * ty) + ty * TILE_LEN) - guess[1];
            }
            n++;
        }
    free(mask);
    return n;
}

/* ---------------------------------------------------- NDT refinement stage -- */

/*
 * Second half of the reference matcher, matchScanCustom.m:32-50: after the grid
 * search, `matchScans(curr, ref, 'InitialPose', pose, 'MaxIterations', 500,
 * 'CellSize', 0.1)` refines the pose continuously and is accepted iff the refined
 * pose passes isValidPose and 2*ndtScore > gridScore; the covariance stays the
 * grid stage's.  matchScans is MathWorks' NDT matcher (not in the reference
 * tree): PARITY UNPINNED.  What follows is OUR restatement of the published
 * algorithm (Biber & Strasser, "The Normal Distributions Transform", IROS 2003):
 *   - reference set = the occupancy set the grid stage scores on (proximity
 *     kernel included), one reference point per set lattice point;
 *   - NDT cells of 0.1 m = 2x2 lattice points, the four grids shifted by half a
 *     cell are the four parities of the 2x2 blocks; a block with >= 3 points
 *     carries a Gaussian (mean, covariance 1/n sum dd^T of its points);
 *   - score S(p) = sum over curr points and the four blocks containing them of
 *     exp(-d^T C^-1 d / 2), p = (tx, ty in cells, phi in rad) applied on top of
 *     the guess; maximised by Newton steps with Levenberg damping on the exact
 *     gradient and Hessian, from the grid stage's optimum, at most 500 iterations
 *     (matchScanCustom.m:36), stop when the gain of an accepted step < 1e-6.
 * Bit-exact contract with the CUDA kernel (k_match.cu mt_ndt_refine): exp/sin/
 * cos are the polynomial forms below (plain IEEE + - * /), per-beam terms are
 * summed in the order 12 warps x 32 lanes, xor-butterfly inside a warp, warps
 * 0-5 and 6-11 sequentially, then the two halves.
 */
#define NDT_MAX_ITERS 500      /* matchScanCustom.m:36 */
#define NDT_TERMS 16         /* S, gradient (3), Hessian (6), curvature model (6) */
#define NDT_SLOTS 384          /* 12 warps x 32 lanes: slot = beam index */
#define NDT_LIMIT 235.0        /* lookups stay inside the occupancy window */
#define NDT_STEP_T 5e-3       /* cells */
#define NDT_STEP_R 5e-5       /* rad */
#define NDT_T1 0.33333333333333331
#define NDT_T2 0.66666666666666663

static int g_refine = 0;
/* 1: orc_match()/orc_match_adj() run the NDT stage after the grid stage. */
int orc_set_refine(int on) { int old = g_refine; g_refine = on; return old; }

/* exp(-e), e >= 0: k = round(-e*log2(e)), Cody-Waite reduction, Taylor to r^13 in
 * Horner form with fused multiply-adds (fma() is correctly rounded on both sides). */
static double ndt_exp_neg(double e)
{
    if (!(e < 700.0)) return 0.0;
    double x = -e;
    double kf = floor(fma(x, 1.4426950408889634, 0.5));
    double r = fma(-kf, 1.90821492927058770002e-10, fma(-kf, 0.693147180369123816490, x));
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    union { int64_t i; double d; } sc;
    sc.i = (int64_t)((int)kf + 1023) << 52;                                  /* 2^k, k >= -1010 */
    return p * sc.d;
}

/* sin and cos of a small angle (|a| < 1): Taylor series in a^2. */
static void ndt_sincos(double a, double *sn, double *cs)
{
    double z = a * a;
    double s = -1.0 / 121645100408832000.0;                                  /* 19! */
    s = s * z + 1.0 / 355687428096000.0;                                     /* 17! */
    s = s * z - 1.0 / 1307674368000.0;                                       /* 15! */
    s = s * z + 1.0 / 6227020800.0;
    s = s * z - 1.0 / 39916800.0;
    s = s * z + 1.0 / 362880.0;
    s = s * z - 1.0 / 5040.0;
    s = s * z + 1.0 / 120.0;
    s = s * z - 1.0 / 6.0;
    s = s * z + 1.0;
    *sn = s * a;
    double c = 1.0 / 2432902008176640000.0;                                  /* 20! */
    c = c * z - 1.0 / 6402373705728000.0;                                    /* 18! */
    c = c * z + 1.0 / 20922789888000.0;                                      /* 16! */
    c = c * z - 1.0 / 87178291200.0;
    c = c * z + 1.0 / 479001600.0;
    c = c * z - 1.0 / 3628800.0;
    c = c * z + 1.0 / 40320.0;
    c = c * z - 1.0 / 720.0;
    c = c * z + 1.0 / 24.0;
    c = c * z - 0.5;
    c = c * z + 1.0;
    *cs = c;
}

/* S, gradient (3) and curvature C (6: 00 01 02 11 12 22) of the NDT score at p.
 * C = sum w G J^T B J is the positive semi-definite part of -Hessian (the
 * majoriser of exp(-e) linearised in e), so every damped step is an ascent
 * direction even where the narrow Gaussians are not concave.
 * win = the (2R+1)^2 occupancy window of match_core, centred on the guess cell.
 * In lattice units (reference points at integers) a curr point (u, v) lies in
 * the four blocks {n, n+1} x {m, m+1}, n in {iu-1, iu}, m in {iv-1, iv}, iu =
 * nearest integer: block extent [n-0.5, n+1.5), so that lattice points are
 * interior.  Each block's Gaussian is weighted by the C1 window w(t) = 1 - 3t^2 +
 * 2t^3, t = |u - (n+0.5)| (times the same in v): the two parities of an axis sum
 * to one, the score is continuous where a point changes blocks. */
static void ndt_eval(const unsigned char *win, int S_, int R_, const double *cx, const double *cy, const int *cj,
                     int M, double fx, double fy, const double *p, double *tot)
{
    static const int orders[5] = {16, 8, 4, 2, 1};
    double (*slot)[NDT_TERMS] = (double (*)[NDT_TERMS])calloc(NDT_SLOTS, sizeof(double[NDT_TERMS]));
    double sn, cs;
    ndt_sincos(p[2], &sn, &cs);
    for (int q = 0; q < M; q++) {
        double *a = slot[cj[q]];
        double X = cs * cx[q] - sn * cy[q], Y = sn * cx[q] + cs * cy[q];
        double X20 = X * 20.0, Y20 = Y * 20.0;
        double u = (X + fx) * 20.0 + p[0], v = (Y + fy) * 20.0 + p[1];
        if (!(fabs(u) < NDT_LIMIT && fabs(v) < NDT_LIMIT)) continue;
        int iu = (int)floor(u + 0.5), iv = (int)floor(v + 0.5);
        double f = 0.0, fu = 0.0, fv = 0.0, fuu = 0.0, fuv = 0.0, fvv = 0.0, cuu = 0.0, cuv = 0.0, cvv = 0.0;
        for (int by = 0; by < 2; by++)
            for (int bx = 0; bx < 2; bx++) {
                int nx = iu - 1 + bx, ny = iv - 1 + by;
                const unsigned char *r0 = win + (size_t)(ny + R_) * S_ + (nx + R_), *r1 = r0 + S_;
                int pat = r0[0] | (r0[1] << 1) | (r1[0] << 2) | (r1[1] << 3);
                int n = (pat & 1) + ((pat >> 1) & 1) + ((pat >> 2) & 1) + (pat >> 3);
                if (n < 3) continue;
                double mx = 0.5, my = 0.5, Bd = 4.0, Bxy = 0.0;
                if (n == 3) {
                    int miss = pat == 14 ? 0 : (pat == 13 ? 1 : (pat == 11 ? 2 : 3));   /* bit = x + 2y of the empty corner */
                    mx = (miss & 1) ? NDT_T1 : NDT_T2;
                    my = (miss >> 1) ? NDT_T1 : NDT_T2;
                    Bd = 6.0;
                    Bxy = ((miss & 1) == (miss >> 1)) ? 3.0 : -3.0;
                }
                double ru = u - (double)nx, rv = v - (double)ny;              /* in [-0.5, 1.5) */
                double tu = ru - 0.5, tv = rv - 0.5;
                double au = fmin(fabs(tu), 1.0), av = fmin(fabs(tv), 1.0);
                double su = tu < 0.0 ? -1.0 : 1.0, sv = tv < 0.0 ? -1.0 : 1.0;
                double wx = fma(-(au * au), fma(-2.0, au, 3.0), 1.0), wy = fma(-(av * av), fma(-2.0, av, 3.0), 1.0);
                double wx1 = su * (6.0 * au * (au - 1.0)), wy1 = sv * (6.0 * av * (av - 1.0));
                double wx2 = fma(12.0, au, -6.0), wy2 = fma(12.0, av, -6.0);
                double dx = ru - mx, dy = rv - my;
                double a1 = fma(Bd, dx, Bxy * dy), a2 = fma(Bxy, dx, Bd * dy);
                double G = ndt_exp_neg(0.5 * fma(dx, a1, dy * a2));
                double W = wx * wy, WG = W * G, Wu = wx1 * wy, Wv = wx * wy1;
                double Gu = -(a1 * G), Gv = -(a2 * G);
                f += WG;
                fu = fma(-a1, WG, fma(Wu, G, fu));
                fv = fma(-a2, WG, fma(Wv, G, fv));
                fuu = fma(fma(a1, a1, -Bd), WG, fma(2.0 * Wu, Gu, fma(wx2 * wy, G, fuu)));
                fuv = fma(fma(a1, a2, -Bxy), WG, fma(Wv, Gu, fma(Wu, Gv, fma(wx1 * wy1, G, fuv))));
                fvv = fma(fma(a2, a2, -Bd), WG, fma(2.0 * Wv, Gv, fma(wx * wy2, G, fvv)));
                cuu = fma(fmax(-wx2, 0.0) * wy, G, fma(WG, Bd, cuu));           /* + concave part of the window */
                cuv = fma(WG, Bxy, cuv);
                cvv = fma(wx * fmax(-wy2, 0.0), G, fma(WG, Bd, cvv));
            }
        /* chain rule to p = (tx, ty, phi): du/dphi = -Y20, dv/dphi = X20, second derivatives -X20, -Y20 */
        double J3x = -Y20, J3y = X20;
        double h13 = fuu * J3x + fuv * J3y, h23 = fuv * J3x + fvv * J3y;
        double c13 = cuu * J3x + cuv * J3y, c23 = cuv * J3x + cvv * J3y;
        a[0] = f;
        a[1] = fu;
        a[2] = fv;
        a[3] = fu * J3x + fv * J3y;
        a[4] = fuu;
        a[5] = fuv;
        a[6] = h13;
        a[7] = fvv;
        a[8] = h23;
        a[9] = (J3x * h13 + J3y * h23) - (fu * X20 + fv * Y20);
        a[10] = cuu;
        a[11] = cuv;
        a[12] = c13;
        a[13] = cvv;
        a[14] = c23;
        a[15] = J3x * c13 + J3y * c23;
    }
    for (int e = 0; e < NDT_TERMS; e++) {
        double half[2] = {0.0, 0.0};                                          /* warps 0-5 and 6-11, each in order */
        for (int w = 0; w < NDT_SLOTS / 32; w++) {
            double l[32], n2[32];
            for (int i = 0; i < 32; i++) l[i] = slot[32 * w + i][e];
            for (int o = 0; o < 5; o++) {
                for (int i = 0; i < 32; i++) n2[i] = l[i] + l[i ^ orders[o]];
                memcpy(l, n2, sizeof l);
            }
            int h = w / (NDT_SLOTS / 64);
            half[h] = (w % (NDT_SLOTS / 64)) ? half[h] + l[0] : l[0];
        }
        tot[e] = half[0] + half[1];
    }
    free(slot);
}

/* Solve A d = g (A symmetric: 00 10 20 11 21 22) by Cholesky with reciprocal pivots;
 * 0 when A is not positive definite. */
static int ndt_solve(double A00, double A10, double A20, double A11, double A21, double A22, const double *g, double *d)
{
    if (!(A00 > 1e-12)) return 0;
    double i00 = 1.0 / sqrt(A00), l10 = A10 * i00, l20 = A20 * i00;
    double t1 = A11 - l10 * l10;
    if (!(t1 > 1e-12)) return 0;
    double i11 = 1.0 / sqrt(t1), l21 = (A21 - l20 * l10) * i11;
    double t2 = (A22 - l20 * l20) - l21 * l21;
    if (!(t2 > 1e-12)) return 0;
    double i22 = 1.0 / sqrt(t2);
    double y0 = g[0] * i00, y1 = (g[1] - l10 * y0) * i11, y2 = ((g[2] - l20 * y0) - l21 * y1) * i22;
    d[2] = y2 * i22;
    d[1] = (y1 - l21 * d[2]) * i11;
    d[0] = ((y0 - l10 * d[1]) - l20 * d[2]) * i00;
    return 1;
}

/* p (in/out) = correction (cells, cells, rad); returns the score at the final p, *evals = score evaluations.
 * Every iteration proposes the Newton step (-H d = g) when -H is positive definite and the step is short
 * (near the optimum: quadratic convergence), otherwise the damped step of the curvature model
 * ((C + lam diag C) d = g); a proposal is kept only if the score increases. */
static double ndt_refine(const unsigned char *win, int S_, int R_, const double *cx, const double *cy, const int *cj,
                         int M, double fx, double fy, double *p, int *evals)
{
    double t[NDT_TERMS], tn[NDT_TERMS], d[3], pn[3], lam = 1e-3;
    ndt_eval(win, S_, R_, cx, cy, cj, M, fx, fy, p, t);
    int ne = 1, newton_ok = 1;
    for (int it = 0; it < NDT_MAX_ITERS; it++) {
        int newton = newton_ok && ndt_solve(-t[4], -t[5], -t[6], -t[7], -t[8], -t[9], t + 1, d) &&
                     fabs(d[0]) < 1.0 && fabs(d[1]) < 1.0 && fabs(d[2]) < 0.01;
        int ok = newton;
        if (!newton)
            ok = ndt_solve(t[10] + lam * t[10], t[11], t[12], t[13] + lam * t[13], t[14], t[15] + lam * t[15], t + 1, d);
        if (!ok && !newton && !(t[10] > 0.0)) break;                           /* no point carries a Gaussian */
        if (ok) {
            if (fabs(d[0]) < NDT_STEP_T && fabs(d[1]) < NDT_STEP_T && fabs(d[2]) < NDT_STEP_R) break;   /* step below 0.25 mm / 5e-5 rad */
            for (int a = 0; a < 3; a++) pn[a] = p[a] + d[a];
            ok = fabs(pn[0]) < 64.0 && fabs(pn[1]) < 64.0 && fabs(pn[2]) < 1.0;
        }
        if (ok) {
            ndt_eval(win, S_, R_, cx, cy, cj, M, fx, fy, pn, tn);
            ne++;
            ok = tn[0] > t[0];
        }
        if (ok) {
            double gain = tn[0] - t[0];
            memcpy(t, tn, sizeof t);
            memcpy(p, pn, sizeof pn);
            if (!newton) lam = lam * 0.1 < 1e-3 ? 1e-3 : lam * 0.1;
            newton_ok = 1;
            if (gain < 1e-6) break;
        } else if (newton) {
            newton_ok = 0;                                                     /* same point again with the model step */
        } else {
            lam *= 10.0;
            if (lam > 1e9) break;
        }
    }
    *evals = ne;
    return t[0];
}

/* Test hook: while set, every match also evaluates the NDT terms at p (cells, cells,
 * rad) into out16 = {S, gradient 3, Hessian 6, curvature model 6}.  Not thread-safe. */
static const double *g_probe_p = NULL;
static double *g_probe_out = NULL;
void orc_ndt_probe(const double *p, double *out16) { g_probe_p = p; g_probe_out = out16; }

/* Shared search core.  Occupancy comes either from the particle's map (m != NULL,
 * scan-to-map, hybridmap.py:210-261) or from the previous scan's endpoints
 * rasterised on the same lattice (ref_x/ref_y, n_ref; scan-to-scan,
 * hybridmap.py:147-191). */
/* The reference's `ref` set (hybridmap.py:230-239): occupied cells are handed to the matcher only when they
 * lie in the 72 x 72-cell window [j - 36, j + 36) of some curr point in some tile (GridMap._get_rel_cell and
 * get_nearby_occ_points, gridmap.py:130-155, float expressions and the dec_x / dec_y quirk replayed) and
 * within 11.5 m of the guess (:239).  Windows come from EVERY curr point of :216-228, also those the
 * 11.0 m filter of :240 drops afterwards.  mask covers global read-lattice cells G0 + [-R1, R1]^2, row b. */
static void match_ref_mask(const orc_map *m, const double *guess, const double *px, const double *py, const double *dist,
                           int B, int G0x, int G0y, int R1, unsigned char *mask)
{
    const int S1 = 2 * R1 + 1, pr = (int)(1.8 / CS);
    memset(mask, 0, (size_t)S1 * S1);
    double c0, s0;
    rb_sincos(guess[2], &s0, &c0);
    for (int j = 0; j < B; j++) {
        if (!(dist[j] < MATCH_MAX_R && dist[j] > MATCH_MIN_R)) continue;      /* :217-218 */
        double gx, gy;
        xform(c0, s0, guess[0], guess[1], px[j], py[j], &gx, &gy);
        orc_tile *t0 = map_with_pos(m, gx, gy);                               /* :220-221 */
        int ix, iy;
        if (!t0 || !get_cell(gx - t0->cx, gy - t0->cy, &ix, &iy)) continue;
        const double cpx = index_to_distance(ix) + t0->cx, cpy = index_to_distance(iy) + t0->cy;   /* :226-228 */
        for (int i = 0; i < m->n; i++) {                                      /* :231-234 every tile */
            const orc_tile *t = &m->tiles[i];
            const double x = cpx - t->cx, y = cpy - t->cy;
            int decx = 0, decy = 0;
            if (y < -TILE_LEN / 2.0) decy = 1; else if (x < -TILE_LEN / 2.0) decx = 1;            /* gridmap.py:131-136 */
            const int jx = (int)(x / TILE_LEN * DIM + DIM / 2.0) - decx, jy = (int)(y / TILE_LEN * DIM + DIM / 2.0) - decy;
            const int x0 = jx - pr > 0 ? jx - pr : 0, y0 = jy - pr > 0 ? jy - pr : 0;
            const int x1 = jx + pr < DIM ? jx + pr : DIM, y1 = jy + pr < DIM ? jy + pr : DIM;
            const int tx = t->cx / TILE_LEN, ty = t->cy / TILE_LEN;
            for (int a = x0; a < x1; a++) {
                const int ra = 800 * tx + a - 400 - G0x;
                if (ra < -R1 || ra > R1) continue;
                const double rx_ = (index_to_distance(a) + t->cx) - guess[0];                     /* :237 */
                for (int b = y0; b < y1; b++) {
                    const int rb = 800 * ty + b - 400 - G0y;
                    if (rb < -R1 || rb > R1) continue;
                    const double ry_ = (index_to_distance(b) + t->cy) - guess[1];
                    if (sqrt(rx_ * rx_ + ry_ * ry_) < MATCH_MAX_R + 0.5) mask[(size_t)(rb + R1) * S1 + (ra + R1)] = 1;   /* :239 */
                }
            }
        }
    }
}

static int match_core(const orc_map *m, const double *ref_x, const double *ref_y, int n_ref,
                      const double *guess, double *cx, double *cy, const int *cj, int M, double rx, double ry,
                      double *out_pose, double *out_cov, double *out_score, int *dbg, int *slice,
                      const double *px, const double *py, const double *dist, int B)
{
    /* cell of the guess position and the guess's offset inside it */
    int t0x = (int)floor(guess[0] / TILE_LEN + 0.5), t0y = (int)floor(guess[1] / TILE_LEN + 0.5);
    if (guess[0] < t0x * TILE_LEN - 20.0) t0x--; else if (guess[0] >= t0x * TILE_LEN + 20.0) t0x++;
    if (guess[1] < t0y * TILE_LEN - 20.0) t0y--; else if (guess[1] >= t0y * TILE_LEN + 20.0) t0y++;
    int i0x = (int)((guess[0] - t0x * TILE_LEN) / TILE_LEN * DIM + DIM / 2.0);
    int i0y = (int)((guess[1] - t0y * TILE_LEN) / TILE_LEN * DIM + DIM / 2.0);
    int G0x = 800 * t0x + i0x - 400, G0y = 800 * t0y + i0y - 400;
    double fx = guess[0] - (index_to_distance(i0x) + t0x * TILE_LEN);
    double fy = guess[1] - (index_to_distance(i0y) + t0y * TILE_LEN);

    int nx = orc_window_cells(rx), ny = orc_window_cells(ry), nk = orc_rot_count();
    double step = orc_rot_step();
    int W = 2 * MAX_NT + 1;
    /* dense occupancy window around the guess cell: every lookup of the search
     * lands inside it (|c| < 11 m = 220 cells, +1 rounding, +14 shift) */
    enum { R = 240, S = 2 * R + 1 };
    unsigned char *win = (unsigned char *)malloc((size_t)S * S);
    unsigned char *raw = (unsigned char *)calloc((size_t)(S + 2) * (S + 2), 1);
    if (m) {
        /* occupied cells of the particle's tiles that the reference would pass on as `ref` (hybridmap.py:230-239) */
        unsigned char *mask = (unsigned char *)malloc((size_t)(S + 2) * (S + 2));
        match_ref_mask(m, guess, px, py, dist, B, G0x, G0y, R + 1, mask);
        for (int b = -R - 1; b <= R + 1; b++)
            for (int a = -R - 1; a <= R + 1; a++) {
                const size_t o = (size_t)(b + R + 1) * (S + 2) + (a + R + 1);
                raw[o] = mask[o] ? (unsigned char)occ_cell(m, G0x + a, G0y + b) : 0;
            }
        free(mask);
    } else {
        /* previous-scan endpoints relative to the guess, |r| < 11 (hybridmap.py:170-171),
         * rasterised like the curr points at rotation 0 */
        for (int q = 0; q < n_ref; q++) {
            double qx = ref_x[q] - guess[0], qy = ref_y[q] - guess[1];
            if (!(sqrt(qx * qx + qy * qy) < MATCH_MAX_R)) continue;
            int a = (int)floor((qx + fx) * 20.0 + 0.5), b = (int)floor((qy + fy) * 20.0 + 0.5);
            if (abs(a) > R || abs(b) > R) continue;
            raw[(size_t)(b + R + 1) * (S + 2) + (a + R + 1)] = 1;
        }
    }
    /* 3x3 proximity kernel: a lookup cell counts when it or one of its 8
     * neighbours is occupied (stands in for the smoothed grid matchScansGrid
     * rasterises from the reference points) */
    for (int b = 0; b < S; b++)
        for (int a = 0; a < S; a++) {
            unsigned char v = 0;
            for (int db = 0; db < 3; db++)
                for (int da = 0; da < 3; da++) v |= raw[(size_t)(b + db) * (S + 2) + (a + da)];
            win[(size_t)b * S + a] = v;
        }
    free(raw);
    int *vol = (int *)calloc((size_t)(2 * nk + 1) * W * W, sizeof(int));
    int *bx = (int *)malloc(sizeof(int) * (size_t)(M ? M : 1)), *by = (int *)malloc(sizeof(int) * (size_t)(M ? M : 1));
    int64_t best = -1; int bi = 0, bj = 0, bk = 0;
    for (int k = -nk; k <= nk; k++) {
        double ck = cos(k * step), sk = sin(k * step);
        int *sl = vol + (size_t)(k + nk) * W * W;
        for (int q = 0; q < M; q++) {
            double rxq = (ck * cx[q] - sk * cy[q]) + fx;
            double ryq = (sk * cx[q] + ck * cy[q]) + fy;
            /* nearest lattice point: the curr points are cell corners (hybridmap.py:226-228), so at
             * rotation 0 rxq*20 is an integer up to rounding noise -- flooring it would make the
             * result depend on the last ulp of the guess */
            bx[q] = (int)floor(rxq * 20.0 + 0.5);      /* cell offsets from the guess cell */
            by[q] = (int)floor(ryq * 20.0 + 0.5);
            if (abs(bx[q]) + nx > R || abs(by[q]) + ny > R) { fprintf(stderr, "orc_match: window overflow\n"); abort(); }
        }
        for (int q = 0; q < M; q++)
            for (int j = -ny; j <= ny; j++) {
                const unsigned char *row = win + (size_t)(by[q] + j + R) * S + (bx[q] + R);
                int *o = sl + (j + MAX_NT) * W + MAX_NT;
                for (int i = -nx; i <= nx; i++) o[i] += row[i];
            }
        for (int j = -ny; j <= ny; j++)
            for (int i = -nx; i <= nx; i++) {
                int64_t key = match_key(sl[(j + MAX_NT) * W + (i + MAX_NT)], i, j, k);
                if (key > best) { best = key; bi = i; bj = j; bk = k; }
            }
    }
    int bs = (int)(best >> 32);
    if (g_probe_p) ndt_eval(win, S, R, cx, cy, cj, M, fx, fy, g_probe_p, g_probe_out);
    if (dbg) { dbg[0] = M; dbg[1] = bi; dbg[2] = bj; dbg[3] = bk; dbg[4] = nx; dbg[5] = ny; dbg[6] = 0; dbg[7] = 0; }
    if (slice)
        for (int j = 0; j < W; j++)
            for (int i = 0; i < W; i++) slice[j * W + i] = vol[((size_t)(bk + nk) * W + j) * W + i];
    out_pose[0] = guess[0] + bi * CS;                                          /* hybridmap.py:253-255 */
    out_pose[1] = guess[1] + bj * CS;
    out_pose[2] = guess[2] + bk * step;
    int valid = fabs(bi * CS) < rx && fabs(bj * CS) < ry && fabs(bk * step) < ROT_RANGE &&
                (bi != 0 || bj != 0 || bk != 0);                                /* matchScanCustom.m:52-57 */
    if (!valid) {
        for (int a = 0; a < 9; a++) out_cov[a] = NAN;                          /* matchScanCustom.m:26-28 */
        *out_score = 0.0;
    } else {
        int64_t W0 = 0, Wx = 0, Wy = 0, Wxx = 0, Wyy = 0, Wxy = 0;
        for (int j = -ny; j <= ny; j++)
            for (int i = -nx; i <= nx; i++) {
                int d = bs - vol[((size_t)(bk + nk) * W + (j + MAX_NT)) * W + (i + MAX_NT)];
                if (d > 40) continue;
                int64_t w = (int64_t)1 << (40 - d);
                W0 += w; Wx += w * i; Wy += w * j; Wxx += w * i * i; Wyy += w * j * j; Wxy += w * i * j;
            }
        int64_t T0 = 0, T1 = 0, T2 = 0;
        int klo = bk - ROT_LINE_HALF < -nk ? -nk : bk - ROT_LINE_HALF;
        int khi = bk + ROT_LINE_HALF > nk ? nk : bk + ROT_LINE_HALF;
        for (int k = klo; k <= khi; k++) {
            int d = bs - vol[((size_t)(k + nk) * W + (bj + MAX_NT)) * W + (bi + MAX_NT)];
            if (d > 40) continue;
            int64_t w = (int64_t)1 << (40 - d);
            T0 += w; T1 += w * k; T2 += w * k * k;
        }
        double mx = (double)Wx / (double)W0, my = (double)Wy / (double)W0, mt = (double)T1 / (double)T0;
        double q = CS * CS, qt = step * step;
        for (int a = 0; a < 9; a++) out_cov[a] = 0.0;
        out_cov[0] = ((double)Wxx / (double)W0 - mx * mx) * q + q / 12.0;
        out_cov[4] = ((double)Wyy / (double)W0 - my * my) * q + q / 12.0;
        out_cov[1] = out_cov[3] = ((double)Wxy / (double)W0 - mx * my) * q;
        out_cov[8] = ((double)T2 / (double)T0 - mt * mt) * qt + qt / 12.0;
        *out_score = (double)bs;
        if (g_refine) {                                                       /* matchScanCustom.m:32-50 */
            double pr[3] = {(double)bi, (double)bj, bk * step};
            int evals = 0;
            double Sn = ndt_refine(win, S, R, cx, cy, cj, M, fx, fy, pr, &evals);
            int ok = fabs(pr[0] * CS) < rx && fabs(pr[1] * CS) < ry && fabs(pr[2]) < ROT_RANGE &&
                     (pr[0] != 0.0 || pr[1] != 0.0 || pr[2] != 0.0);           /* :38 isValidPose(new_pose) */
            int accepted = ok && Sn * 2.0 > (double)bs;                         /* :39 */
            if (accepted) {
                out_pose[0] = guess[0] + pr[0] * CS;
                out_pose[1] = guess[1] + pr[1] * CS;
                out_pose[2] = guess[2] + pr[2];
                *out_score = Sn;
            }
            if (dbg) { dbg[6] = evals; dbg[7] = accepted; }
        }
    }
    free(win); free(vol); free(bx); free(by);
    return valid;
}

int orc_match(const orc_map *m, const double *guess, const double *px, const double *py,
              const double *dist, int B, double rx, double ry,
              double *out_pose, double *out_cov, double *out_score, int *dbg, int *slice)
{
    double *cx = (double *)malloc(sizeof(double) * (size_t)B), *cy = (double *)malloc(sizeof(double) * (size_t)B);
    int *cj = (int *)malloc(sizeof(int) * (size_t)B);
    int M = match_curr(m, guess, px, py, dist, B, cx, cy, cj);
    int v = match_core(m, NULL, NULL, 0, guess, cx, cy, cj, M, rx, ry, out_pose, out_cov, out_score, dbg, slice, px, py, dist, B);
    free(cx); free(cy); free(cj);
    return v;
}

/* Scan-to-previous-scan variant, HybridMap.get_scan_adj hybridmap.py:147-191:
 * curr = every beam endpoint at the guess minus the guess position (not snapped,
 * no range gate), ref = previous scan's global endpoints minus the guess
 * position, both kept when |.| < 11.0 (:171-172); same search contract. */
int orc_match_adj(const double *guess, const double *px, const double *py, int B,
                  const double *prev_x, const double *prev_y, int n_prev, double rx, double ry,
                  double *out_pose, double *out_cov, double *out_score, int *dbg, int *slice)
{
    double *cx = (double *)malloc(sizeof(double) * (size_t)B), *cy = (double *)malloc(sizeof(double) * (size_t)B);
    int *cj = (int *)malloc(sizeof(int) * (size_t)B);
    double c0, s0;
    rb_sincos(guess[2], &s0, &c0);
    int M = 0;
    for (int j = 0; j < B; j++) {
        double gx, gy;
        xform(c0, s0, guess[0], guess[1], px[j], py[j], &gx, &gy);
        double qx = gx - guess[0], qy = gy - guess[1];
        if (!(sqrt(qx * qx + qy * qy) < MATCH_MAX_R)) continue;
        cx[M] = qx; cy[M] = qy; cj[M] = j; M++;
    }
    int v = match_core(NULL, prev_x, prev_y, n_prev, guess, cx, cy, cj, M, rx, ry, out_pose, out_cov, out_score, dbg, slice, px, py, NULL, B);
    free(cx); free(cy); free(cj);
    return v;
}

/* -------------------------------------------------------------- motion -- */

/* Robot.imu_update robot.py:45-57 with the three loader families:
 *   0 absolute-set       IntelIMUData.py:22-36   u = (x, y, theta)
 *   1 additive velocity  IntelRawIMUData.py:33-55 (same shape: Aces, Freid*, Obero, Bele)
 *                        u = (vx, vy, w), par = (a_xy, b_xy, a_th, b_th)
 *   2 unicycle           DefaultIMUData.py:25-54  u = (v, w)
 * dt is in seconds (the loaders divide the 1e-4 s tick count by 1e4). */
static void mat3mul(const double *a, const double *b, double *o)
{
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            o[3 * r + c] = (a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c]) + a[3 * r + 2] * b[6 + c];
}

void orc_motion(int family, const double *u, double dt, const double *par, double *pose, double *cov)
{
    double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Q[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, np_[3];
    if (family == 0) {
        np_[0] = u[0]; np_[1] = u[1]; np_[2] = u[2];
        F[0] = 0.01 * 0.01; F[4] = 0.01 * 0.01; F[8] = (0.2 * M_PI / 180) * (0.2 * M_PI / 180);
        Q[0] = Q[4] = Q[8] = 1.0;
        Q[2] = u[0] - pose[0];
        Q[5] = u[1] - pose[1];
    } else if (family == 1) {
        np_[0] = pose[0] + u[0] * dt; np_[1] = pose[1] + u[1] * dt; np_[2] = pose[2] + u[2] * dt;
        double q0 = par[0] + par[1] * fabs(u[0]) * dt, q1 = par[0] + par[1] * fabs(u[1]) * dt,
               q2 = par[2] + par[3] * fabs(u[2]) * dt;
        Q[0] = fabs(q0 * q0); Q[4] = fabs(q1 * q1); Q[8] = fabs(q2 * q2);
    } else {
        double th = pose[2] + dt * u[1];
        double ct, st, cp, sp;
        rb_sincos(th, &st, &ct);
        rb_sincos(pose[2], &sp, &cp);
        np_[0] = pose[0] + dt * u[0] * ct;
        np_[1] = pose[1] + dt * u[0] * st;
        np_[2] = th;
        F[2] = dt * u[0] * cp;
        F[5] = dt * u[0] * sp;
        double g0 = dt * cp, g1 = dt * sp, g2 = dt, m0 = 0.05 * 0.05, m1 = (M_PI / 180 / 2) * (M_PI / 180 / 2);
        /* |G M G^T| + noise, DefaultIMUData.py:44-54 ; G = [[g0,0],[g1,0],[0,g2]] */
        Q[0] = fabs((g0 * m0) * g0) + 0.01 * 0.01;
        Q[1] = fabs((g0 * m0) * g1);
        Q[3] = fabs((g1 * m0) * g0);
        Q[4] = fabs((g1 * m0) * g1) + 0.01 * 0.01;
        Q[8] = fabs((g2 * m1) * g2) + (0.2 * M_PI / 180) * (0.2 * M_PI / 180);
    }
    double t1[9], Ft[9], t2[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Ft[3 * r + c] = F[3 * c + r];
    mat3mul(F, cov, t1);                                    /* robot.py:50 */
    mat3mul(t1, Ft, t2);
    for (int a = 0; a < 9; a++) cov[a] = t2[a] + Q[a];      /* robot.py:51 */
    pose[0] = np_[0]; pose[1] = np_[1]; pose[2] = np_[2];
}

/* ------------------------------------------------------------ resample -- */

/* main.resample main.py:46-79 on float64 weights (declared pin).
 * Returns 0 = no resample (ancestors = identity), 1 = resampled, -1 = the
 * reference's AssertionError("Incorrect number of resampled weights."). */
int orc_resample(const double *weights, int N, double u01, int *anc)
{
    double mx = weights[0], mn = weights[0];
    for (int i = 1; i < N; i++) { if (weights[i] > mx) mx = weights[i]; if (weights[i] < mn) mn = weights[i]; }
    for (int i = 0; i < N; i++) anc[i] = i;
    if (!(mx - mn > RESAMPLE_TRIGGER)) return 0;                          /* :50 */
    double *w = (double *)malloc(sizeof(double) * (size_t)N);
    for (int i = 0; i < N; i++) w[i] = weights[i] == -INFINITY ? 0.0 : weights[i];   /* :53 */
    mn = w[0];
    for (int i = 1; i < N; i++) if (w[i] < mn) mn = w[i];
    if (mn < 0) { double a = fabs(mn); for (int i = 0; i < N; i++) if (w[i] != 0) w[i] += a; }  /* :54-55 */
    double tot = 0.0;
    for (int i = 0; i < N; i++) tot = tot + w[i];                          /* :57 builtin sum, left to right */
    double slice = tot / N;
    double start = u01 * slice;                                            /* :59 */
    double cur = 0.0;
    long emitted = 0;
    int ok = 1;
    for (int i = 0; i < N; i++) {                                          /* :61-64 */
        cur += w[i];
        double f = floor((cur - start) / slice);
        if (!(f > -4e18 && f < 4e18)) { ok = 0; break; }                  /* math.floor raises on nan/inf */
        long num = (long)f - emitted + 1;
        for (long q = 0; q < num; q++) { if (emitted < N) anc[emitted] = i; emitted++; }
    }
    free(w);
    if (!ok || emitted != N) return -1;                                    /* :66-67 */
    return 1;
}

/* -------------------------------------------------------- whole filter -- */

#ifdef _OPENMP
#include <omp.h>
#endif
/* Number of OpenMP threads of the particle loops below (torchrun exports
 * OMP_NUM_THREADS=1, which would silently make the CPU baseline single-threaded).
 * n <= 0 only queries.  Returns the thread count in effect. */
int orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    return 1;
#endif
}

typedef struct {
    int N, B, K;
    double *pose;      /* N*3 */
    double *cov;       /* N*9 */
    double *weight;    /* N   current weight = weight()[-1] */
    orc_map **map;     /* N   */
    int *valid;        /* N   last match validity */
    double *px, *py, *dist;
} orc_filter;

/* particles = [Robot(eng) ...] main.py:87 ; Robot.__init__ robot.py:20-28. */
orc_filter *orc_filter_new(int N, int B, int K)
{
    orc_filter *f = (orc_filter *)calloc(1, sizeof(orc_filter));
    f->N = N; f->B = B; f->K = K;
    f->pose = (double *)calloc((size_t)N * 3, sizeof(double));
    f->cov = (double *)calloc((size_t)N * 9, sizeof(double));
    f->weight = (double *)malloc(sizeof(double) * (size_t)N);
    f->map = (orc_map **)malloc(sizeof(orc_map *) * (size_t)N);
    f->valid = (int *)calloc((size_t)N, sizeof(int));
    f->px = (double *)calloc((size_t)B, sizeof(double));
    f->py = (double *)calloc((size_t)B, sizeof(double));
    f->dist = (double *)calloc((size_t)B, sizeof(double));
    for (int i = 0; i < N; i++) { f->weight[i] = 1.0; f->map[i] = orc_map_new(); }
    return f;
}

void orc_filter_free(orc_filter *f)
{
    if (!f) return;
    for (int i = 0; i < f->N; i++) orc_map_free(f->map[i]);
    free(f->pose); free(f->cov); free(f->weight); free(f->map); free(f->valid);
    free(f->px); free(f->py); free(f->dist); free(f);
}

double *orc_filter_pose(orc_filter *f) { return f->pose; }
double *orc_filter_cov(orc_filter *f) { return f->cov; }
double *orc_filter_weight(orc_filter *f) { return f->weight; }
int *orc_filter_valid(orc_filter *f) { return f->valid; }
orc_map *orc_filter_map(orc_filter *f, int i) { return f->map[i]; }

void orc_filter_set_scan(orc_filter *f, const double *ranges, const double *angles)
{
    orc_scan_prepare(ranges, angles, f->B, f->px, f->py, f->dist);
}

/* [p.imu_update(reading) for p in particles] main.py:144. */
void orc_filter_motion(orc_filter *f, int family, const double *u, double dt, const double *par)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < f->N; i++) orc_motion(family, u, dt, par, f->pose + 3 * i, f->cov + 9 * i);
}

/* Integrate the current scan at every particle's pose (the commented-out map
 * seeding at main.py:89-90). */
void orc_filter_integrate(orc_filter *f)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < f->N; i++) orc_map_update(f->map[i], f->pose + 3 * i, f->px, f->py, f->dist, f->B);
}

/* One particle of Robot.map_update robot.py:59-115 (scan-to-map branch), z = K*3 normals. */
/* z_is_guesses: z holds the K proposal samples themselves (drawn by the caller with NumPy, robot.py:81)
 * instead of standard normals for our mean + chol(cov) z transform. */
static void particle_map_update(orc_filter *f, int i, const double *z, const double *prev_x, const double *prev_y,
                                int n_prev, int z_is_guesses)
{
    double *pose = f->pose + 3 * i, *cov = f->cov + 9 * i;
    orc_map *m = f->map[i];
    double rx, ry, mp[3], mc[9], sc;
    orc_pose_range(cov, &rx, &ry);
    int valid = prev_x ? orc_match_adj(pose, f->px, f->py, f->B, prev_x, prev_y, n_prev, rx, ry, mp, mc, &sc, NULL, NULL)  /* robot.py:66-67 */
                       : orc_match(m, pose, f->px, f->py, f->dist, f->B, rx, ry, mp, mc, &sc, NULL, NULL);             /* robot.py:68-69 */
    f->valid[i] = valid;
    if (!valid) {                                                          /* :73-78 */
        double one = 1.0, w;
        orc_map_update(m, pose, f->px, f->py, f->dist, f->B);
        orc_sample_weight(m, pose, 1, f->px, f->py, f->dist, f->B, &one, &w);
        f->weight[i] = w + f->weight[i];
        return;
    }
    int K = f->K;
    double *g = (double *)malloc(sizeof(double) * (size_t)K * 5), *prs = g + 3 * K, *w = g + 4 * K;
    double mean[3], sigma[9];
    if (z_is_guesses) {                                                    /* :80-87 */
        memcpy(g, z, sizeof(double) * 3 * (size_t)K);
        orc_propose_pdf(mp, mc, g, K, prs);
    } else {
        orc_propose(mp, mc, z, K, g, prs);
    }
    orc_sample_weight(m, g, K, f->px, f->py, f->dist, f->B, prs, w);       /* :88 */
    double norm = orc_moments(g, w, K, mean, sigma);                       /* :89-108 */
    for (int a = 0; a < 3; a++) pose[a] = mean[a];                         /* :109-113 */
    for (int a = 0; a < 9; a++) cov[a] = sigma[a];
    f->weight[i] = norm + f->weight[i];                                    /* :114 */
    orc_map_update(m, pose, f->px, f->py, f->dist, f->B);                  /* :115 */
    free(g);
}

/* [p.map_update(scan, last_scan, False) for p in particles] main.py:157.
 * z holds N*K*3 standard normals, particle-major. */
void orc_filter_map_update(orc_filter *f, const double *z)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < f->N; i++) particle_map_update(f, i, z + (size_t)3 * f->K * i, NULL, NULL, 0, 0);
}

/* [p.map_update(scan, last_scan, True) for p in particles] main.py:159 : every
 * particle matches against the same previous scan (global endpoints, main.py:168). */
void orc_filter_map_update_adj(orc_filter *f, const double *z, const double *prev_x, const double *prev_y, int n_prev)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < f->N; i++) particle_map_update(f, i, z + (size_t)3 * f->K * i, prev_x, prev_y, n_prev, 0);
}

/* The same two calls with the proposal samples supplied (N*K*3, rows of particles whose match fails are
 * ignored): the caller drew them with np.random.multivariate_normal from the matcher result, robot.py:81. */
void orc_filter_map_update_guesses(orc_filter *f, const double *g, const double *prev_x, const double *prev_y, int n_prev)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < f->N; i++) particle_map_update(f, i, g + (size_t)3 * f->K * i, prev_x, prev_y, n_prev, 1);
}

/* particles = resample(particles) main.py:160 with Robot.copy robot.py:141-149. */
int orc_filter_resample(orc_filter *f, double u01, int *anc)
{
    int N = f->N, rc = orc_resample(f->weight, N, u01, anc);
    if (rc != 1) return rc;
    double *pose = (double *)malloc(sizeof(double) * (size_t)N * 3), *cov = (double *)malloc(sizeof(double) * (size_t)N * 9);
    orc_map **map = (orc_map **)malloc(sizeof(orc_map *) * (size_t)N);
    char *used = (char *)calloc((size_t)N, 1);
    for (int j = 0; j < N; j++) {                                          /* :69-75 */
        int a = anc[j];
        memcpy(pose + 3 * j, f->pose + 3 * a, 3 * sizeof(double));
        memcpy(cov + 9 * j, f->cov + 9 * a, 9 * sizeof(double));
        if (!used[a]) { map[j] = f->map[a]; used[a] = 1; } else map[j] = NULL;
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < N; j++) if (!map[j]) map[j] = orc_map_copy(f->map[anc[j]]);
    for (int a = 0; a < N; a++) if (!used[a]) orc_map_free(f->map[a]);
    memcpy(f->pose, pose, sizeof(double) * (size_t)N * 3);
    memcpy(f->cov, cov, sizeof(double) * (size_t)N * 9);
    memcpy(f->map, map, sizeof(orc_map *) * (size_t)N);
    for (int j = 0; j < N; j++) f->weight[j] = 1.0;                        /* :77-78 */
    free(pose); free(cov); free(map); free(used);
    return 1;
}
