/*
 * amc_oracle.c -- CPU restatement of the Argon_Monte_Carlo per-timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under argon_monte_carlo_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker.
 *
 * Every function cites the reference lines it restates (paths relative to the upstream
 * repository: Cube = Open_Air_Cube_MC.py, Pore = Open_Air_Pore_MC.py,
 * Temp = Temperature_Pore_MC.py).  The restatement is pinned against the Python reference
 * itself (oracle/make_golden.py imports it under a matplotlib stub) and against the shipped
 * momentum_energy.csv; see oracle/README.md.
 *
 * Two arithmetic modes:
 *   ref mode   (orc_set_ref_mode(1)): reproduces CPython/NumPy scalar arithmetic on this
 *              image bit for bit -- NumPy-scalar `v**2` is libm pow(v, 2.0) and np.dot on
 *              2/3-vectors is OpenBLAS ddot whose scalar tail is an FMA chain.
 *   plain mode (default): `v**2` is v*v and dot products are unfused left-to-right sums.
 *              This is the arithmetic the CUDA path implements, so GPU-vs-oracle tests are
 *              bit-exact in this mode.  The two modes differ by <= 1 ulp per operation.
 * Everything else (operation order, separate mul/add roundings, case order, index order) is
 * identical in both modes.  Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ arithmetic modes */
static int g_ref_mode = 0;
void orc_set_ref_mode(int on) { g_ref_mode = on; }
int orc_get_ref_mode(void) { return g_ref_mode; }

static inline double sq(double v) { return g_ref_mode ? pow(v, 2.0) : v * v; }
static inline double dot2(double a0, double a1, double b0, double b1)
{
    if (g_ref_mode) return fma(a1, b1, a0 * b0);
    return a0 * b0 + a1 * b1;
}
static inline double dot3(double a0, double a1, double a2, double b0, double b1, double b2)
{
    if (g_ref_mode) return fma(a2, b2, fma(a1, b1, a0 * b0));
    return (a0 * b0 + a1 * b1) + a2 * b2;
}

/* ------------------------------------------------------------------ shared structs */
typedef struct {
    int64_t n;
    double *x, *y, *z, *vx, *vy, *vz;
    double *dist, *dist_x, *dist_y, *dist_z; /* dist_since_collision + components */
    uint8_t *flag;                           /* full_path_traveled */
    double *px, *py, *pz;                    /* prior_{x,y,z}_vals */
} orc_state;

/* completed_{,x_,y_,z_}paths: four parallel append-only lists (Pore:410-413) */
typedef struct {
    int64_t count, cap;
    double *total, *cx, *cy, *cz;
} orc_paths;

/* collision-pair log (not materialised by the reference; test tap) */
typedef struct {
    int64_t count, cap;
    int64_t *hi, *lo; /* larger / smaller global particle index of the pair */
    int32_t *group;   /* colour group 0..7 (Cube sweep: 0) */
    int32_t *cell;    /* linear reference cell id */
} orc_pairlog;

static void paths_push(orc_paths *s, double t, double a, double b, double c)
{
    if (!s) return;
    if (s->count == s->cap) {
        int64_t nc = s->cap ? s->cap * 2 : 1024;
        s->total = (double *)realloc(s->total, nc * sizeof(double));
        s->cx = (double *)realloc(s->cx, nc * sizeof(double));
        s->cy = (double *)realloc(s->cy, nc * sizeof(double));
        s->cz = (double *)realloc(s->cz, nc * sizeof(double));
        s->cap = nc;
    }
    s->total[s->count] = t;
    s->cx[s->count] = a;
    s->cy[s->count] = b;
    s->cz[s->count] = c;
    s->count++;
}
void orc_paths_free(orc_paths *s)
{
    free(s->total); free(s->cx); free(s->cy); free(s->cz);
    memset(s, 0, sizeof(*s));
}
static void pairlog_push(orc_pairlog *l, int64_t hi, int64_t lo, int group, int cell)
{
    if (!l) return;
    if (l->count == l->cap) {
        int64_t nc = l->cap ? l->cap * 2 : 1024;
        l->hi = (int64_t *)realloc(l->hi, nc * sizeof(int64_t));
        l->lo = (int64_t *)realloc(l->lo, nc * sizeof(int64_t));
        l->group = (int32_t *)realloc(l->group, nc * sizeof(int32_t));
        l->cell = (int32_t *)realloc(l->cell, nc * sizeof(int32_t));
        l->cap = nc;
    }
    l->hi[l->count] = hi; l->lo[l->count] = lo;
    l->group[l->count] = group; l->cell[l->count] = cell;
    l->count++;
}
void orc_pairlog_free(orc_pairlog *l)
{
    free(l->hi); free(l->lo); free(l->group); free(l->cell);
    memset(l, 0, sizeof(*l));
}

/* ------------------------------------------------------------------ drift
 * Pore:427-437, Temp:673-683, Cube:180-187.  `x += dt*v` is a rounded multiply followed by
 * a rounded add; np.square on arrays is an exact-rounded multiply in both modes. */
void orc_drift(orc_state *s, double dt, int save_prior)
{
    for (int64_t i = 0; i < s->n; i++) {
        if (save_prior) { s->px[i] = s->x[i]; s->py[i] = s->y[i]; s->pz[i] = s->z[i]; }
        double ax = dt * s->vx[i], ay = dt * s->vy[i], az = dt * s->vz[i];
        s->x[i] += ax; s->y[i] += ay; s->z[i] += az;
        s->dist[i] += fabs(sqrt((ax * ax + ay * ay) + az * az));
        s->dist_x[i] += fabs(ax);
        s->dist_y[i] += fabs(ay);
        s->dist_z[i] += fabs(az);
    }
}

/* ------------------------------------------------------------------ MFP bookkeeping helper
 * The pattern shared by Pore:274-284, 324-335 and the pair loop Pore:186-199: if the
 * particle already completed a first collision, append |path - |speed*t|| (and component
 * analogues) to the four lists, else set its flag. */
static void mfp_record(orc_state *s, int64_t i, double vx, double vy, double vz, double t,
                       orc_paths *sink)
{
    if (s->flag[i]) {
        double sp = sqrt((sq(vx) + sq(vy)) + sq(vz));
        paths_push(sink, fabs(s->dist[i] - fabs(sp * t)), fabs(s->dist_x[i] - fabs(vx * t)),
                   fabs(s->dist_y[i] - fabs(vy * t)), fabs(s->dist_z[i] - fabs(vz * t)));
    } else {
        s->flag[i] = 1;
    }
}

/* ------------------------------------------------------------------ pore / temperature geometry
 * All thresholds are evaluated on the host with the reference's own Python expressions
 * (e.g. open_air_height + hot_coating_height = 1.3000000000000003e-07) and passed in. */
typedef struct {
    double argon_mass, argon_radius, collision_range;
    double R_oa, R_oa_c;   /* open_air_radius, open_air_collision_radius   Pore:35,67 */
    double R_p, R_p_c;     /* pore_coated_radius, pore_collision_radius    Pore:25,69 */
    double R_g, R_g_c;     /* gap_radius, gap_collision_radius             Pore:26,68 */
    double H;              /* total_height                                  Pore:39 */
    double oah;            /* open_air_height                               Pore:36 */
    double z_cold;         /* total_height - open_air_height                Pore:457 */
    double z_gb;           /* open_air_height + hot_coating_height          Pore:465 (gap_bottom_height Temp:45) */
    double z_gt_pore;      /* total_height-open_air_height-cold_coating_height  Pore:465 */
    double z_gt;           /* open_air_height+hot_coating_height+gap_height Pore:371 (gap_top_height Temp:46) */
    double ten_a;          /* 10*argon_radius                               Pore:358 */
    double R_oa_sq, R_g_sq, R_p_sq; /* Python-float `**2` of the radii      Pore:363,367,371 */
    /* Temp-only thresholds */
    double zc3;            /* total_height - open_air_height + argon_radius Temp:708 */
    double zh3;            /* open_air_height - argon_radius                Temp:713 */
    double zgt_m;          /* gap_top_height - argon_radius                 Temp:720 */
    double zgb_p;          /* gap_bottom_height + argon_radius              Temp:720 */
    double R_g_c_sq, R_p_c_sq; /* collision radii squared (np.float64**2 -> pow)  Temp:721,728 */
    double recap_lo, recap_hi; /* 50e-9, total_height - 50e-9               Temp:599,602 */
    double E_cold, E_hot;  /* surface_energy_cold / _hot                    Temp:83-84 */
    double alpha_c, alpha_g; /* accommodation coefficients                  Temp:76-77 */
    double cos85;          /* cos(85*pi/180)                                Temp:136 */
} orc_geom;

/* ---- specular cylinder side wall with MFP bookkeeping.  Pore:294-348.
 * Returns 1 on a floating-point error (the reference's try/except path): particle untouched. */
static int side_wall_quadratic(double x, double y, double vx, double vy, double Rc, double *t_out)
{
    double a = sq(-vx) + sq(-vy);
    double b = 2 * (x * (-vx) + y * (-vy));
    double c = (sq(x) + sq(y)) - sq(Rc);
    double disc = sq(b) - (4 * a) * c;
    if (!(disc >= 0.0) || a == 0.0) return 1; /* invalid sqrt / divide by zero raise under seterr */
    double r = sqrt(disc);
    double t1 = (-b + r) / (2 * a), t2 = (-b - r) / (2 * a);
    *t_out = t1 < t2 ? t1 : t2; /* np.min */
    return 0;
}

static int pore_side_wall_one(orc_state *s, int64_t i, double Rc, orc_paths *sink)
{
    double x = s->x[i], y = s->y[i], vx = s->vx[i], vy = s->vy[i], vz = s->vz[i];
    double t;
    if (side_wall_quadratic(x, y, vx, vy, Rc, &t)) return 1;
    double col_x = x - vx * t, col_y = y - vy * t;
    double nx = col_x / Rc, ny = col_y / Rc;
    double scalar = dot2(vx, vy, nx, ny);
    double nvx = vx - (2 * scalar) * nx, nvy = vy - (2 * scalar) * ny;
    double new_x = col_x + nvx * t, new_y = col_y + nvy * t;
    mfp_record(s, i, vx, vy, vz, t, sink);
    s->x[i] = new_x; s->y[i] = new_y; s->vx[i] = nvx; s->vy[i] = nvy;
    s->dist[i] = fabs(sqrt((sq(nvx) + sq(nvy)) + sq(vz)) * t);
    s->dist_x[i] = fabs(nvx * t);
    s->dist_y[i] = fabs(nvy * t);
    s->dist_z[i] = fabs(vz * t);
    return 0;
}

/* ---- specular plane wall with MFP bookkeeping.  Pore:257-292. */
static void pore_plane_wall_one(orc_state *s, int64_t i, double zp, orc_paths *sink)
{
    double vx = s->vx[i], vy = s->vy[i], vz = s->vz[i];
    double t = (s->z[i] - zp) / vz;
    mfp_record(s, i, vx, vy, vz, t, sink);
    s->dist[i] = fabs(sqrt((sq(vx) + sq(vy)) + sq(vz)) * t);
    s->dist_x[i] = fabs(vx * t);
    s->dist_y[i] = fabs(vy * t);
    s->dist_z[i] = fabs(vz * t);
    s->vz[i] = -vz;
    s->z[i] = zp + t * s->vz[i];
}

/* ---- Pore wall cases 1..6, evaluated sequentially against the mutated state (Pore:442-485).
 * counts[9] = hits per case in the order 1, 2a, 2b, 3cold, 3hot, 4, 5bottom, 5top, 6.
 * hit_bits (nullable): per particle, bit k set when case k hit.  Returns number of FP errors. */
int64_t orc_pore_walls(orc_state *s, const orc_geom *g, orc_paths *sink, uint16_t *hit_bits,
                       int64_t *counts)
{
    int64_t errs = 0, n = s->n;
    for (int k = 0; k < 9; k++) counts[k] = 0;
    if (hit_bits) memset(hit_bits, 0, n * sizeof(uint16_t));
#define HIT(k) do { counts[k]++; if (hit_bits) hit_bits[i] |= (uint16_t)(1u << (k)); } while (0)
    /* case 1  Pore:442-443 */
    for (int64_t i = 0; i < n; i++)
        if (sqrt(s->x[i] * s->x[i] + s->y[i] * s->y[i]) > g->R_oa) { HIT(0); errs += pore_side_wall_one(s, i, g->R_oa_c, sink); }
    /* case 2a Pore:448-449 */
    for (int64_t i = 0; i < n; i++)
        if (s->z[i] < 0) { HIT(1); pore_plane_wall_one(s, i, 0.0, sink); }
    /* case 2b Pore:451-452 */
    for (int64_t i = 0; i < n; i++)
        if (s->z[i] > g->H) { HIT(2); pore_plane_wall_one(s, i, g->H, sink); }
    /* case 3 cold Pore:457-458 */
    for (int64_t i = 0; i < n; i++)
        if (s->pz[i] > g->z_cold && s->z[i] < g->z_cold && sqrt(s->x[i] * s->x[i] + s->y[i] * s->y[i]) > g->R_p) { HIT(3); pore_plane_wall_one(s, i, g->z_cold, sink); }
    /* case 3 hot Pore:460-461 */
    for (int64_t i = 0; i < n; i++)
        if (s->pz[i] < g->oah && s->z[i] > g->oah && sqrt(s->x[i] * s->x[i] + s->y[i] * s->y[i]) > g->R_p) { HIT(4); pore_plane_wall_one(s, i, g->oah, sink); }
    /* case 4 Pore:465-467 */
    for (int64_t i = 0; i < n; i++)
        if (s->pz[i] < g->z_gt_pore && s->pz[i] > g->z_gb && sqrt(s->px[i] * s->px[i] + s->py[i] * s->py[i]) < g->R_g &&
            sqrt(s->x[i] * s->x[i] + s->y[i] * s->y[i]) > g->R_g) { HIT(5); errs += pore_side_wall_one(s, i, g->R_g_c, sink); }
    /* case 5 bottom Pore:472-474 */
    for (int64_t i = 0; i < n; i++)
        if (sqrt(s->px[i] * s->px[i] + s->py[i] * s->py[i]) > g->R_p && s->z[i] < g->z_gb && s->pz[i] < g->z_gt_pore && s->pz[i] > g->z_gb) { HIT(6); pore_plane_wall_one(s, i, g->z_gb, sink); }
    /* case 5 top Pore:476-478 */
    for (int64_t i = 0; i < n; i++)
        if (sqrt(s->px[i] * s->px[i] + s->py[i] * s->py[i]) > g->R_p && s->z[i] > g->z_gt_pore && s->pz[i] < g->z_gt_pore && s->pz[i] > g->z_gb) { HIT(7); pore_plane_wall_one(s, i, g->z_gt_pore, sink); }
    /* case 6 Pore:482-485 */
    for (int64_t i = 0; i < n; i++) {
        double r = sqrt(s->x[i] * s->x[i] + s->y[i] * s->y[i]);
        if (sqrt(s->px[i] * s->px[i] + s->py[i] * s->py[i]) < g->R_p && r > g->R_p &&
            ((s->z[i] < g->z_cold && s->z[i] > g->z_gt_pore) || (s->z[i] < g->z_gb && s->z[i] > g->oah))) { HIT(8); errs += pore_side_wall_one(s, i, g->R_p_c, sink); }
    }
#undef HIT
    return errs;
}

/* ---- Pore num_out_of_bounds(): counts AND teleports.  Pore:354-375. */
int64_t orc_pore_recapture(orc_state *s, const orc_geom *g)
{
    int64_t cnt = 0, n = s->n;
    for (int64_t i = 0; i < n; i++) if (s->z[i] < 0) { s->z[i] += g->ten_a; cnt++; }
    for (int64_t i = 0; i < n; i++) if (s->z[i] > g->H) { s->z[i] -= g->ten_a; cnt++; }
    for (int64_t i = 0; i < n; i++) if (s->x[i] * s->x[i] + s->y[i] * s->y[i] > g->R_oa_sq) { s->x[i] = 0; s->y[i] = 0; cnt++; }
    for (int64_t i = 0; i < n; i++)
        if (s->x[i] * s->x[i] + s->y[i] * s->y[i] > g->R_g_sq && s->z[i] > g->oah && s->z[i] < g->z_cold) { s->x[i] = 0; s->y[i] = 0; cnt++; }
    for (int64_t i = 0; i < n; i++)
        if (s->x[i] * s->x[i] + s->y[i] * s->y[i] > g->R_p_sq &&
            ((s->z[i] > g->oah && s->z[i] < g->z_gb) || (s->z[i] > g->z_gt && s->z[i] < g->z_cold))) { s->x[i] = 0; s->y[i] = 0; cnt++; }
    return cnt;
}

/* ------------------------------------------------------------------ Temp walls
 * Case ids (order of evaluation, Temp:693-753):
 *   0: case 1 (specular open-air side)      1: 2a (z<0)          2: 2b (z>H)
 *   3: 3 cold plane   4: 3 hot plane   5: 4 gap side   6: 5 bottom (hot)   7: 5 top (cold)
 *   8: 6 hot side     9: 6 cold side
 * The energized cases consume host RNG in the reference (Temp:132-141), so they are split into
 * detect (mask -> ascending index list + the `norm` argument random_inbounds_direction gets)
 * and apply (given one accepted unit direction per valid hit). */
enum { TC_1 = 0, TC_2A, TC_2B, TC_3C, TC_3H, TC_4, TC_5B, TC_5T, TC_6H, TC_6C, TC_COUNT };

static int temp_mask(const orc_state *s, const orc_geom *g, int c, int64_t i)
{
    double x = s->x[i], y = s->y[i], z = s->z[i], px = s->px[i], py = s->py[i], pz = s->pz[i];
    switch (c) {
    case TC_1:  return sqrt(x * x + y * y) > g->R_oa;                                   /* Temp:693 */
    case TC_2A: return z < 0;                                                            /* Temp:699 */
    case TC_2B: return z > g->H;                                                         /* Temp:702 */
    case TC_3C: return pz >= g->zc3 && z < g->zc3 && x * x + y * y > g->R_p_sq;          /* Temp:708 */
    case TC_3H: return pz <= g->zh3 && z > g->zh3 && x * x + y * y > g->R_p_sq;          /* Temp:713 */
    case TC_4:  return pz < g->zgt_m && pz > g->zgb_p && px * px + py * py <= g->R_g_c_sq && x * x + y * y > g->R_g_c_sq; /* Temp:720-721 */
    case TC_5B: return px * px + py * py >= g->R_p_c_sq && z < g->zgb_p && pz <= g->zgt_m && pz >= g->zgb_p; /* Temp:728-729 */
    case TC_5T: return px * px + py * py >= g->R_p_c_sq && z > g->zgt_m && pz <= g->zgt_m && pz >= g->zgb_p; /* Temp:734-735 */
    case TC_6H: return px * px + py * py <= g->R_p_c_sq && x * x + y * y > g->R_p_c_sq && z <= g->zgb_p && z >= g->zh3; /* Temp:743-744 */
    case TC_6C: return px * px + py * py <= g->R_p_c_sq && x * x + y * y > g->R_p_c_sq && z < g->zc3 && z > g->zgt_m;   /* Temp:749-750 */
    }
    return 0;
}

static double temp_case_plane(const orc_geom *g, int c)
{
    switch (c) {
    case TC_2A: return 0.0;
    case TC_2B: return g->H;
    case TC_3C: return g->zc3;
    case TC_3H: return g->zh3;
    case TC_5B: return g->zgb_p;
    case TC_5T: return g->zgt_m;
    }
    return 0.0;
}
static double temp_case_inbound(int c) /* inbounds_direction argument, Temp:709,714,730,736 */
{
    return (c == TC_3C || c == TC_5B) ? 1.0 : -1.0;
}
static double temp_case_radius(const orc_geom *g, int c)
{
    if (c == TC_1) return g->R_oa_c;
    if (c == TC_4) return g->R_g_c;
    return g->R_p_c;
}
static int temp_case_is_side(int c) { return c == TC_1 || c == TC_4 || c == TC_6H || c == TC_6C; }

/* detect: fills idx[] (ascending), normal[3*k] (the vector handed to random_inbounds_direction,
 * NaN when the hit raises a floating-point error and therefore draws nothing), colz[k]
 * (contact height; surface_energy_gap's argument for case 4).  Returns the hit count. */
int64_t orc_temp_case_detect(const orc_state *s, const orc_geom *g, int c, int64_t *idx, double *normal,
                             double *colz)
{
    int64_t k = 0;
    for (int64_t i = 0; i < s->n; i++) {
        if (!temp_mask(s, g, c, i)) continue;
        if (idx) idx[k] = i;
        if (normal) {
            double nx = 0, ny = 0, nz = 0, cz = 0;
            if (temp_case_is_side(c)) {
                double t;
                if (side_wall_quadratic(s->x[i], s->y[i], s->vx[i], s->vy[i], temp_case_radius(g, c), &t)) {
                    nx = ny = nz = NAN; cz = NAN;
                } else {
                    double Rc = temp_case_radius(g, c);
                    double col_x = s->x[i] - s->vx[i] * t, col_y = s->y[i] - s->vy[i] * t;
                    nx = -(col_x / Rc); ny = -(col_y / Rc); nz = -(0.0 / Rc); /* -normalized_norm_vect Temp:442-444 */
                    cz = s->z[i] - s->vz[i] * t;
                }
            } else {
                nz = temp_case_inbound(c); cz = temp_case_plane(g, c);
            }
            normal[3 * k] = nx; normal[3 * k + 1] = ny; normal[3 * k + 2] = nz;
            if (colz) colz[k] = cz;
        }
        k++;
    }
    return k;
}

/* shared energized-exchange arithmetic, Temp:377-389 / 446-458 / 516-527.  The reference runs
 * this partly in mpmath mpf at 53 bits, round-to-nearest: same results as IEEE double. */
static void energized_exchange(const orc_geom *g, double vx, double vy, double vz, double Es, double alpha,
                               const double *dir, double *nv, double *dpz, double *dE)
{
    double v_mag = sqrt((sq(vx) + sq(vy)) + sq(vz));
    double old_pz = g->argon_mass * vz;
    double E = (0.5 * g->argon_mass) * sq(v_mag);
    double ediff = Es - E;
    double Enew = E + ediff * alpha;
    double new_mag = sqrt((Enew * 2) / g->argon_mass);
    *dE = Enew - E;
    nv[0] = dir[0] * new_mag; nv[1] = dir[1] * new_mag; nv[2] = dir[2] * new_mag;
    double new_pz = g->argon_mass * nv[2];
    *dpz = new_pz - old_pz;
}

/* apply one case.  dirs[3*k]: accepted unit direction for hit k (ignored for specular cases and
 * for error hits); surf_e[k]: surface energy per hit for case 4 (surface_energy_gap(col_z),
 * Temp:519), ignored otherwise.  sums[0] += sum of dpz, sums[1] += sum of dE, accumulated
 * sequentially in ascending index order like the reference.  Returns number of FP errors. */
int64_t orc_temp_case_apply(orc_state *s, const orc_geom *g, int c, int64_t nh, const int64_t *idx,
                            const double *dirs, const double *surf_e, orc_paths *sink, double *sums)
{
    int64_t errs = 0;
    double sum_p = 0.0, sum_e = 0.0; /* Python int 0 + mpf: exact */
    for (int64_t k = 0; k < nh; k++) {
        int64_t i = idx[k];
        if (c == TC_2A || c == TC_2B) { /* hit_vertical_specular_wall Temp:311-315 */
            double zp = temp_case_plane(g, c);
            double t = (s->z[i] - zp) / s->vz[i];
            s->vz[i] = -s->vz[i];
            s->z[i] = zp + t * s->vz[i];
            continue;
        }
        if (c == TC_1) { /* hit_cylinder_specular_side_wall Temp:317-347 */
            double x = s->x[i], y = s->y[i], vx = s->vx[i], vy = s->vy[i], t, Rc = g->R_oa_c;
            if (side_wall_quadratic(x, y, vx, vy, Rc, &t)) { errs++; continue; }
            double col_x = x - vx * t, col_y = y - vy * t;
            double nx = col_x / Rc, ny = col_y / Rc;
            double scalar = dot2(vx, vy, nx, ny);
            double nvx = vx - (2 * scalar) * nx, nvy = vy - (2 * scalar) * ny;
            s->x[i] = col_x + nvx * t; s->y[i] = col_y + nvy * t; s->vx[i] = nvx; s->vy[i] = nvy;
            continue;
        }
        double x = s->x[i], y = s->y[i], z = s->z[i], vx = s->vx[i], vy = s->vy[i], vz = s->vz[i];
        double t, col_x, col_y, col_z, Es, alpha = g->alpha_c;
        if (temp_case_is_side(c)) { /* Temp:414-483 / 485-553 */
            if (side_wall_quadratic(x, y, vx, vy, temp_case_radius(g, c), &t)) { errs++; continue; }
            col_x = x - vx * t; col_y = y - vy * t; col_z = z - vz * t;
        } else { /* hit_vertical_coated_wall Temp:349-412 */
            double zp = temp_case_plane(g, c);
            t = (z - zp) / vz;
            col_x = x - vx * t; col_y = y - vy * t; col_z = zp;
        }
        if (c == TC_4) { Es = surf_e[k]; alpha = g->alpha_g; }
        else Es = (c == TC_3C || c == TC_5T || c == TC_6C) ? g->E_cold : g->E_hot;
        double nv[3], dpz, dE;
        energized_exchange(g, vx, vy, vz, Es, alpha, dirs + 3 * k, nv, &dpz, &dE);
        sum_p += dpz;
        if (c != TC_4) sum_e += dE;
        mfp_record(s, i, vx, vy, vz, t, sink);
        s->dist[i] = 0; s->dist_x[i] = 0; s->dist_y[i] = 0; s->dist_z[i] = 0;
        s->x[i] = col_x; s->y[i] = col_y; s->z[i] = col_z;
        s->vx[i] = nv[0]; s->vy[i] = nv[1]; s->vz[i] = nv[2];
    }
    if (sums) { sums[0] = sum_p; sums[1] = sum_e; }
    return errs;
}

/* Temp recapture_out_of_bounds().  Temp:594-616. */
int64_t orc_temp_recapture(orc_state *s, const orc_geom *g)
{
    int64_t cnt = 0, n = s->n;
    for (int64_t i = 0; i < n; i++) if (s->z[i] < 0) { s->z[i] = g->recap_lo; cnt++; }
    for (int64_t i = 0; i < n; i++) if (s->z[i] > g->H) { s->z[i] = g->recap_hi; cnt++; }
    for (int64_t i = 0; i < n; i++) if (s->x[i] * s->x[i] + s->y[i] * s->y[i] > g->R_oa_sq) { s->x[i] = 0; s->y[i] = 0; cnt++; }
    for (int64_t i = 0; i < n; i++)
        if (s->x[i] * s->x[i] + s->y[i] * s->y[i] > g->R_g_sq && s->z[i] > g->oah && s->z[i] < g->z_cold) { s->x[i] = 0; s->y[i] = 0; cnt++; }
    for (int64_t i = 0; i < n; i++)
        if (s->x[i] * s->x[i] + s->y[i] * s->y[i] > g->R_p_sq &&
            ((s->z[i] > g->oah && s->z[i] < g->z_gb) || (s->z[i] > g->z_gt && s->z[i] < g->z_cold))) { s->x[i] = 0; s->y[i] = 0; cnt++; }
    return cnt;
}

/* Temp num_out_of_bounds(): report only (the coordinate prints are not restated).  Temp:560-592. */
int64_t orc_temp_oob_count(const orc_state *s, const orc_geom *g)
{
    int64_t cnt = 0;
    for (int64_t i = 0; i < s->n; i++) {
        double x = s->x[i], y = s->y[i], z = s->z[i], r2 = x * x + y * y;
        cnt += z < 0;
        cnt += z > g->H;
        cnt += r2 > g->R_oa_sq && z >= 0 && z <= g->oah;
        cnt += r2 > g->R_oa_sq && z >= g->z_cold && z <= g->H;
        cnt += r2 > g->R_g_sq && z >= g->z_gb && z <= g->z_gt;
        cnt += r2 > g->R_p_sq && z > g->oah && z < g->z_gb;
        cnt += r2 > g->R_p_sq && z > g->z_gt && z < g->z_cold;
    }
    return cnt;
}

/* ------------------------------------------------------------------ device-RNG (throughput) mode
 * Not in the reference: a counter-based replacement for the two host Mersenne-Twister streams
 * (Temp:119-141).  Philox4x32-10 keyed by the run seed, counter = (particle id, step, case,
 * attempt); an isotropic unit vector by Marsaglia's disc method (only + - * sqrt, so the CUDA
 * path reproduces it bit for bit); then the reference's own acceptance rule Temp:136-139. */
static inline void philox_round(uint32_t *c, uint32_t k0, uint32_t k1)
{
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
static inline double u53(uint32_t hi, uint32_t lo)
{
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}
/* direction about `norm`, deterministic in (seed, id, step, case). */
void orc_philox_direction(uint64_t seed, int64_t id, int64_t step, int c, const double *norm, double cos85,
                          double *dir)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (uint32_t attempt = 0;; attempt++) {
        uint32_t ctr[4] = {(uint32_t)id, (uint32_t)((uint64_t)id >> 32) ^ ((uint32_t)c << 24), (uint32_t)step, attempt};
        uint32_t r[4];
        orc_philox4x32_10(ctr, key, r);
        double u = 2.0 * u53(r[0], r[1]) - 1.0, v = 2.0 * u53(r[2], r[3]) - 1.0;
        double ss = u * u + v * v;
        if (!(ss < 1.0) || ss == 0.0) continue;
        double root = sqrt(1.0 - ss);
        double d0 = (2.0 * u) * root, d1 = (2.0 * v) * root, d2 = 1.0 - 2.0 * ss;
        double dn = (d0 * norm[0] + d1 * norm[1]) + d2 * norm[2];
        if (fabs(dn) < cos85) continue;
        if (dn < cos85) { d0 = -d0; d1 = -d1; d2 = -d2; }
        dir[0] = d0; dir[1] = d1; dir[2] = d2;
        return;
    }
}
/* Chebyshev evaluation of surface_energy_gap(z) on [zlo, zhi]; coefficients fitted on the host
 * with mpmath (Temp:143-152).  Clenshaw recurrence, unfused. */
double orc_cheb_eval(const double *coef, int ncoef, double zmid, double inv_half, double z)
{
    double u = (z - zmid) * inv_half, u2 = 2.0 * u, b1 = 0.0, b2 = 0.0;
    for (int k = ncoef - 1; k >= 1; k--) {
        double b0 = (coef[k] + u2 * b1) - b2;
        b2 = b1; b1 = b0;
    }
    return (coef[0] + u * b1) - b2;
}

/* all ten Temp wall cases with the device-RNG rule.  sums[3] = {dpz, dE_cold, dE_hot} of this
 * step, accumulated case-major in ascending index order.  counts[10] hits per case. */
int64_t orc_temp_walls_philox(orc_state *s, const orc_geom *g, uint64_t seed, int64_t step, const double *coef,
                              int ncoef, double zmid, double inv_half, orc_paths *sink, uint16_t *hit_bits,
                              int64_t *counts, double *sums)
{
    int64_t errs = 0, n = s->n;
    int64_t *idx = (int64_t *)malloc((n ? n : 1) * sizeof(int64_t));
    double *normal = (double *)malloc((n ? n : 1) * 3 * sizeof(double));
    double *colz = (double *)malloc((n ? n : 1) * sizeof(double));
    double *dirs = (double *)malloc((n ? n : 1) * 3 * sizeof(double));
    double *se = (double *)malloc((n ? n : 1) * sizeof(double));
    if (hit_bits) memset(hit_bits, 0, n * sizeof(uint16_t));
    sums[0] = sums[1] = sums[2] = 0.0;
    for (int c = 0; c < TC_COUNT; c++) {
        int64_t nh = orc_temp_case_detect(s, g, c, idx, normal, colz);
        counts[c] = nh;
        for (int64_t k = 0; k < nh; k++) {
            if (hit_bits) hit_bits[idx[k]] |= (uint16_t)(1u << c);
            if (c >= TC_3C && normal[3 * k] == normal[3 * k]) {
                orc_philox_direction(seed, idx[k], step, c, normal + 3 * k, g->cos85, dirs + 3 * k);
                if (c == TC_4) se[k] = orc_cheb_eval(coef, ncoef, zmid, inv_half, colz[k]);
            }
        }
        double cs[2];
        errs += orc_temp_case_apply(s, g, c, nh, idx, dirs, se, sink, cs);
        if (c >= TC_3C) {
            sums[0] += cs[0];
            if (c == TC_3C || c == TC_5T || c == TC_6C) sums[1] += cs[1];
            else if (c != TC_4) sums[2] += cs[1];
        }
    }
    free(idx); free(normal); free(colz); free(dirs); free(se);
    return errs;
}

/* ------------------------------------------------------------------ particle-particle pass
 * Cell grid: axis a has nc[a] cells with first cell index c0[a]; edge[a][k] = (c0+k)*d is the
 * reference's `(…)*dx` product and lo[a][k] = edge[a][k] - band its `(…)*dx - collision_range`
 * (Pore:527-529) -- both tables are computed on the host with those Python expressions. */
typedef struct {
    int32_t nc[3];
    int32_t c0[3];
    const double *edge[3]; /* nc+1 entries */
    const double *lo[3];   /* nc entries   */
} orc_grid;

/* The reference's pairwise_particles_in_cell (Pore:160-255; Temp:215-309; Cube:253-324) on the
 * members m[0..nm) of one cell, members in ascending global index.  Operates in place on the
 * global arrays: within one colour group a particle belongs to at most one cell, so this equals
 * the reference's copy-out / scatter-back (Pore:533-547). */
static int64_t pairwise_in_cell(orc_state *s, const int64_t *m, int64_t nm, double cr, double mass,
                                orc_paths *sink, orc_pairlog *plog, int group, int cell, int64_t *errs)
{
    int64_t ncol = 0;
    double guard = cr * 1.000001; /* quick reject only; never decides a borderline pair */
    for (int64_t ii = 0; ii < nm; ii++) {
        int64_t i = m[ii];
        for (int64_t jj = 0; jj < ii; jj++) {
            int64_t j = m[jj];
            double x1 = s->x[j], x2 = s->x[i];
            double ddx = x2 - x1;
            if (fabs(ddx) > guard) continue;
            double y1 = s->y[j], y2 = s->y[i], z1 = s->z[j], z2 = s->z[i];
            double ddy = y2 - y1, ddz = z2 - z1;
            if (fabs(ddy) > guard || fabs(ddz) > guard) continue;
            double sep = sqrt((sq(ddx) + sq(ddy)) + sq(ddz));  /* Pore:173 */
            if (!(sep < cr)) continue;                          /* Pore:174 */
            double vx1 = s->vx[j], vx2 = s->vx[i], vy1 = s->vy[j], vy2 = s->vy[i], vz1 = s->vz[j], vz2 = s->vz[i];
            double rx = -vx2 + vx1, ry = -vy2 + vy1, rz = -vz2 + vz1;
            double a = (sq(rx) + sq(ry)) + sq(rz);                                 /* Pore:182 */
            double b = 2 * ((ddx * rx + ddy * ry) + ddz * rz);                     /* Pore:183 */
            double c = ((sq(ddx) + sq(ddy)) + sq(ddz)) - sq(cr);                   /* Pore:184 */
            double disc = sq(b) - (4 * a) * c;
            if (!(disc >= 0.0) || a == 0.0) { if (errs) (*errs)++; continue; }     /* would raise in the reference */
            double root = sqrt(disc);
            double t1 = (-b + root) / (2 * a), t2 = (-b - root) / (2 * a);
            double t = t1 > t2 ? t1 : t2;                                          /* np.max Pore:185 */
            mfp_record(s, j, vx1, vy1, vz1, t, sink);                              /* Pore:186-192 */
            mfp_record(s, i, vx2, vy2, vz2, t, sink);                              /* Pore:193-199 */
            double nx1 = x1 - vx1 * t, ny1 = y1 - vy1 * t, nz1 = z1 - vz1 * t;     /* Pore:202 */
            double nx2 = x2 - vx2 * t, ny2 = y2 - vy2 * t, nz2 = z2 - vz2 * t;
            double n0 = (nx2 - nx1) / cr, n1 = (ny2 - ny1) / cr, n2 = (nz2 - nz1) / cr; /* Pore:205-207 */
            double p = (dot3(vx1, vy1, vz1, n0, n1, n2) - dot3(vx2, vy2, vz2, n0, n1, n2)) / mass; /* Pore:209 */
            double pm = p * mass;
            double wx1 = vx1 - pm * n0, wy1 = vy1 - pm * n1, wz1 = vz1 - pm * n2;  /* Pore:211-213 */
            double wx2 = vx2 + pm * n0, wy2 = vy2 + pm * n1, wz2 = vz2 + pm * n2;  /* Pore:214-216 */
            s->x[j] = nx1 + wx1 * t; s->y[j] = ny1 + wy1 * t; s->z[j] = nz1 + wz1 * t; /* Pore:218,221-223 */
            s->x[i] = nx2 + wx2 * t; s->y[i] = ny2 + wy2 * t; s->z[i] = nz2 + wz2 * t; /* Pore:219,224-226 */
            s->vx[j] = wx1; s->vy[j] = wy1; s->vz[j] = wz1;
            s->vx[i] = wx2; s->vy[i] = wy2; s->vz[i] = wz2;
            s->dist[i] = fabs(sqrt((sq(wx2) + sq(wy2)) + sq(wz2)) * t);            /* Pore:233 */
            s->dist[j] = fabs(sqrt((sq(wx1) + sq(wy1)) + sq(wz1)) * t);            /* Pore:234 */
            s->dist_x[i] = fabs(wx2 * t); s->dist_y[i] = fabs(wy2 * t); s->dist_z[i] = fabs(wz2 * t);
            s->dist_x[j] = fabs(wx1 * t); s->dist_y[j] = fabs(wy1 * t); s->dist_z[j] = fabs(wz1 * t);
            pairlog_push(plog, i, j, group, cell);
            ncol++;
        }
    }
    return ncol;
}

/* member cell of coordinate v on one axis for the given parity (colour group bit), or -1.
 * Cell k (0-based) of parity `par` contains v iff lo[k] < v < edge[k+1], strict both sides. */
static inline int axis_cell(const double *edge, const double *lo, int nc, int par, double v)
{
    /* find owner: largest k with edge[k] <= v (binary search; tables are tiny) */
    if (!(v > lo[0]) || !(v < edge[nc])) return -1;
    int a = 0, b = nc; /* invariant: edge[a] <= v or a==0 ; v < edge[b] */
    while (b - a > 1) { int mid = (a + b) >> 1; if (edge[mid] <= v) a = mid; else b = mid; }
    /* a is the owner when v >= edge[0]; when lo[0] < v < edge[0], a == 0 and only the band applies */
    for (int k = a; k <= a + 1 && k < nc; k++)
        if ((k & 1) == par && lo[k] < v && v < edge[k + 1]) return k;
    return -1;
}

/* One full particle-particle pass in the 8-colour-group semantics (Pore:522-549, Temp:815-842).
 * Groups run in the order x_group, y_group, z_group (z fastest); membership is re-evaluated
 * from the current positions at the start of every group; cells of a group are independent.
 * checks (nullable): reference-equivalent pair tests, sum over visited cells of n(n-1)/2. */
int64_t orc_pp_groups(orc_state *s, const orc_grid *g, double cr, double mass, orc_paths *sink,
                      orc_pairlog *plog, int64_t *checks, int64_t *errs_out)
{
    int64_t n = s->n, total = 0, nchecks = 0, errs = 0;
    int nhx = (g->nc[0] + 1) / 2, nhy = (g->nc[1] + 1) / 2, nhz = (g->nc[2] + 1) / 2;
    int64_t ncell = (int64_t)nhx * nhy * nhz;
    int32_t *cell_of = (int32_t *)malloc((n ? n : 1) * sizeof(int32_t));
    int64_t *start = (int64_t *)malloc((ncell + 1) * sizeof(int64_t));
    int64_t *members = (int64_t *)malloc((n ? n : 1) * sizeof(int64_t));
    for (int xg = 0; xg < 2; xg++) for (int yg = 0; yg < 2; yg++) for (int zg = 0; zg < 2; zg++) {
        int group = (xg << 2) | (yg << 1) | zg;
        memset(start, 0, (ncell + 1) * sizeof(int64_t));
        for (int64_t i = 0; i < n; i++) {
            int cx = axis_cell(g->edge[0], g->lo[0], g->nc[0], xg, s->x[i]);
            int cy = cx < 0 ? -1 : axis_cell(g->edge[1], g->lo[1], g->nc[1], yg, s->y[i]);
            int cz = cy < 0 ? -1 : axis_cell(g->edge[2], g->lo[2], g->nc[2], zg, s->z[i]);
            if (cz < 0) { cell_of[i] = -1; continue; }
            int32_t c = (int32_t)(((int64_t)(cx >> 1) * nhy + (cy >> 1)) * nhz + (cz >> 1)); /* x-major, z-minor Pore:530 */
            cell_of[i] = c;
            start[c + 1]++;
        }
        for (int64_t c = 0; c < ncell; c++) start[c + 1] += start[c];
        { /* stable fill -> members of each cell in ascending global index (boolean-mask gather order) */
            int64_t *fill = (int64_t *)malloc((ncell ? ncell : 1) * sizeof(int64_t));
            memcpy(fill, start, ncell * sizeof(int64_t));
            for (int64_t i = 0; i < n; i++) if (cell_of[i] >= 0) members[fill[cell_of[i]]++] = i;
            free(fill);
        }
        /* cells are independent inside a group: run them in parallel, keep per-cell outputs and
         * append them in cell order so the result does not depend on the thread count */
        orc_paths *cpaths = sink ? (orc_paths *)calloc(ncell, sizeof(orc_paths)) : NULL;
        orc_pairlog *clog = plog ? (orc_pairlog *)calloc(ncell, sizeof(orc_pairlog)) : NULL;
        int64_t gtotal = 0, gchecks = 0, gerrs = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : gtotal, gchecks, gerrs)
        for (int64_t c = 0; c < ncell; c++) {
            int64_t nm = start[c + 1] - start[c];
            if (nm < 1) continue;
            gchecks += nm * (nm - 1) / 2;
            int64_t e = 0;
            gtotal += pairwise_in_cell(s, members + start[c], nm, cr, mass, cpaths ? &cpaths[c] : NULL,
                                       clog ? &clog[c] : NULL, group, (int)c, &e);
            gerrs += e;
        }
        for (int64_t c = 0; c < ncell; c++) {
            if (cpaths) {
                for (int64_t k = 0; k < cpaths[c].count; k++)
                    paths_push(sink, cpaths[c].total[k], cpaths[c].cx[k], cpaths[c].cy[k], cpaths[c].cz[k]);
                orc_paths_free(&cpaths[c]);
            }
            if (clog) {
                for (int64_t k = 0; k < clog[c].count; k++)
                    pairlog_push(plog, clog[c].hi[k], clog[c].lo[k], clog[c].group[k], clog[c].cell[k]);
                orc_pairlog_free(&clog[c]);
            }
        }
        free(cpaths); free(clog);
        total += gtotal; nchecks += gchecks; errs += gerrs;
    }
    free(cell_of); free(start); free(members);
    if (checks) *checks = nchecks;
    if (errs_out) *errs_out = errs;
    return total;
}

/* ------------------------------------------------------------------ Cube
 * Six specular plane walls, no MFP bookkeeping.  Cube:192-226. */
void orc_cube_walls(orc_state *s, double cube_x, double cube_y, double cube_z, int64_t *counts)
{
    double *pos[3] = {s->x, s->y, s->z}, *vel[3] = {s->vx, s->vy, s->vz};
    double L[3] = {cube_x, cube_y, cube_z};
    for (int a = 0; a < 3; a++) {
        int64_t hi = 0, lo = 0;
        for (int64_t i = 0; i < s->n; i++)
            if (pos[a][i] > L[a]) { double t = (pos[a][i] - L[a]) / vel[a][i]; vel[a][i] = -vel[a][i]; pos[a][i] = L[a] + t * vel[a][i]; hi++; }
        for (int64_t i = 0; i < s->n; i++)
            if (pos[a][i] < 0) { double t = pos[a][i] / vel[a][i]; vel[a][i] = -vel[a][i]; pos[a][i] = t * vel[a][i]; lo++; }
        if (counts) { counts[2 * a] = hi; counts[2 * a + 1] = lo; }
    }
}

/* Serial lexicographic sweep over the cells with write-back after every cell (Cube:232-336).
 * The x-layer mask is taken once per x layer, the y-layer mask once per (x, y) column and the
 * z-layer mask per cell, each from the positions current at that moment (Cube:233,235,237). */
int64_t orc_cube_pp_sweep(orc_state *s, const orc_grid *g, double cr, double mass, orc_paths *sink,
                          orc_pairlog *plog, int64_t *checks, int64_t *errs_out)
{
    int64_t n = s->n, total = 0, nchecks = 0, errs = 0;
    int64_t *lx = (int64_t *)malloc((n ? n : 1) * sizeof(int64_t));
    int64_t *lxy = (int64_t *)malloc((n ? n : 1) * sizeof(int64_t));
    int64_t *cellm = (int64_t *)malloc((n ? n : 1) * sizeof(int64_t));
    for (int xl = 0; xl < g->nc[0]; xl++) {
        int64_t nx = 0;
        for (int64_t i = 0; i < n; i++)
            if (g->lo[0][xl] < s->x[i] && s->x[i] < g->edge[0][xl + 1]) lx[nx++] = i;
        for (int yl = 0; yl < g->nc[1]; yl++) {
            int64_t nxy = 0;
            for (int64_t k = 0; k < nx; k++) {
                int64_t i = lx[k];
                if (g->lo[1][yl] < s->y[i] && s->y[i] < g->edge[1][yl + 1]) lxy[nxy++] = i;
            }
            for (int zl = 0; zl < g->nc[2]; zl++) {
                int64_t nm = 0;
                for (int64_t k = 0; k < nxy; k++) {
                    int64_t i = lxy[k];
                    if (g->lo[2][zl] < s->z[i] && s->z[i] < g->edge[2][zl + 1]) cellm[nm++] = i;
                }
                nchecks += nm * (nm - 1) / 2;
                int cell = (xl * g->nc[1] + yl) * g->nc[2] + zl;
                total += pairwise_in_cell(s, cellm, nm, cr, mass, sink, plog, 0, cell, &errs);
            }
        }
    }
    free(lx); free(lxy); free(cellm);
    if (checks) *checks = nchecks;
    if (errs_out) *errs_out = errs;
    return total;
}

/* ------------------------------------------------------------------ histogram
 * np.histogram(values, bins=200, range=(0, 1e-6)) uniform-bin path with its edge corrections
 * (numpy/lib/_histograms_impl.py:816-873), as called through Axes.hist at Pore:575. */
void orc_histogram(const double *v, int64_t n, int nbins, double first, double last, const double *edges,
                   int64_t *counts)
{
    double norm_numerator = (double)nbins, norm_denom = last - first;
    for (int b = 0; b < nbins; b++) counts[b] = 0;
    for (int64_t i = 0; i < n; i++) {
        double x = v[i];
        if (!(x >= first && x <= last)) continue;
        double f = ((x - first) / norm_denom) * norm_numerator;
        int64_t idx = (int64_t)f; /* astype(np.intp) truncation */
        if (idx == nbins) idx--;
        if (x < edges[idx]) idx--;
        else if (x >= edges[idx + 1] && idx != nbins - 1) idx++;
        counts[idx]++;
    }
}

/* pin the OpenMP team size (bench.py: the same thread count at every GPU count, whatever OMP_NUM_THREADS the launcher exported) */
void orc_set_num_threads(int n)
{
    if (n > 0) omp_set_num_threads(n);
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
