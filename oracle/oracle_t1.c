/* oracle/oracle_t1.c -- TEST INFRASTRUCTURE ONLY: the "T1" CPU oracle (plain C99).
 *
 * Restates, function by function, the reference's exciton hop path.  Citations are into /root/reference/src:
 *
 *   t1_linspace ................. arma::linspace as used at monte_carlo/monte_carlo.cpp:157-167
 *   t1_forster_table ............ monte_carlo::create_forster_scatt_table   monte_carlo/monte_carlo.cpp:156-200
 *   argmin_abs / t1_get_rate .... scattering_struct::get_rate               monte_carlo/scattering_struct.h:38-52
 *   t1_set_mesh ................. monte_carlo::create_scatterers            monte_carlo/monte_carlo.h:199-271
 *   t1_trim ..................... monte_carlo::trim_scats                   monte_carlo/monte_carlo.h:722-782
 *   t1_find_domain .............. monte_carlo::find_simulation_domain       monte_carlo/monte_carlo.h:328-340
 *   t1_build_buckets ............ monte_carlo::create_scatterer_buckets     monte_carlo/monte_carlo.h:368-416
 *   find_neighbors .............. scatterer::find_neighbors                 monte_carlo/scatterer.cpp:34-83
 *   t1_set_max_rate ............. scatterer::set_max_rate + driver          monte_carlo/scatterer.h:89-93, monte_carlo.h:426-440
 *   t1_injection ................ injection_region, get_removal_domain      monte_carlo/monte_carlo.cpp:203-251
 *   t1_select / update_state .... scatterer::update_state                   monte_carlo/scatterer.cpp:9-31
 *   ff_time ..................... scatterer::ff_time                        monte_carlo/scatterer.h:74-80
 *   fly / particle_step ......... particle::fly, particle::step             monte_carlo/particle.cpp:9-54, 57-80
 *   t1_kubo_create_particles .... kubo_create_particles + particle ctor     monte_carlo/monte_carlo.cpp:308-316, particle.h:48-52
 *   t1_kubo_step ................ kubo_step + update_delta_pos + MSD row     monte_carlo/monte_carlo.cpp:319-342, 382-409, particle.h:97
 *   contacts .................... init/create_particles/step/repopulate/...  monte_carlo/monte_carlo.h:157-195, 274-316, 343-355, 443-516, 525-688
 *
 * Third-party arithmetic that is NOT under /root/reference: Armadillo (un-vendored, version unpinned).  Its small-
 * vector reductions are restated from its published algorithms: dot and norm use two partial sums over even and odd
 * indices, added at the end: (x0*y0 + x2*y2) + x1*y1;  index_min is the first strict minimum starting from +inf;
 * normalise divides by the norm, or by 1 if the norm is not > 0.  glibc supplies rand()/log()/acos()/pow()/cos()/sin().
 *
 * Draw sources.  GLIBC: one global sequential stream, rand(), excitons visited in index order -- the reference's
 * OMP_NUM_THREADS=1 behaviour.  PHILOX: draw k of exciton g is word (k&1) of Philox2x32-10(ctr={k>>1, g lo}, key =
 * seed lo ^ rot16(seed hi) ^ g hi * 0x9E3779B9) shifted right by one bit (31 bits, so RAND_MAX arithmetic is unchanged)
 * -- the stream the CUDA engine uses: two words per call, which is what one scattering event consumes.
 * REPLAY: per-exciton lists of recorded draws.
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off (no FMA contraction, like the reference's x86-64 baseline build).
 */
#define _DEFAULT_SOURCE
#include "oracle_t1.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define T1_RAND_MAX 2147483647 /* glibc RAND_MAX */

/* ------------------------------------------------------------------------------------------------------------ */
/* small 3-vector helpers with Armadillo's accumulation order                                                  */
/* ------------------------------------------------------------------------------------------------------------ */
static double dot3(const double a[3], const double b[3]) {
  double v1 = 0.0, v2 = 0.0;
  v1 += a[0] * b[0];
  v2 += a[1] * b[1];
  v1 += a[2] * b[2];
  return v1 + v2;
}
static double norm3(const double a[3]) {
  double v1 = 0.0, v2 = 0.0;
  v1 += a[0] * a[0];
  v2 += a[1] * a[1];
  v1 += a[2] * a[2];
  return sqrt(v1 + v2); /* Armadillo's rescaling fallback only triggers for a zero / non-finite result */
}

/* ------------------------------------------------------------------------------------------------------------ */
/* stand-alone pieces                                                                                          */
/* ------------------------------------------------------------------------------------------------------------ */
void t1_linspace(double start, double end, int64_t n, double* out) {
  if (n >= 2) {
    const double delta = (end - start) / (double)(n - 1);
    for (int64_t i = 0; i < n - 1; ++i) out[i] = start + (double)i * delta;
    out[n - 1] = end;
  } else if (n == 1) {
    out[0] = end;
  }
}

void t1_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Philox2x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) */
void t1_philox2x32_10(const uint32_t ctr[2], uint32_t key, uint32_t out[2]) {
  uint32_t c0 = ctr[0], c1 = ctr[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p = (uint64_t)0xD256D193u * c0;
    const uint32_t n0 = (uint32_t)(p >> 32) ^ key ^ c1;
    c1 = (uint32_t)p;
    c0 = n0;
    key += 0x9E3779B9u;
  }
  out[0] = c0;
  out[1] = c1;
}

int32_t t1_philox_draw(uint64_t seed, uint64_t exciton, uint64_t k) {
  const uint32_t sh = (uint32_t)(seed >> 32), gh = (uint32_t)(exciton >> 32);
  const uint32_t key = (uint32_t)seed ^ ((sh << 16) | (sh >> 16)) ^ (gh * 0x9E3779B9u);
  const uint32_t ctr[2] = {(uint32_t)(k >> 1), (uint32_t)exciton};
  uint32_t       o[2];
  t1_philox2x32_10(ctr, key, o);
  return (int32_t)(o[k & 1] >> 1);
}

/* monte_carlo.cpp:172-193.  normalise(r1) is the zero vector when ash1 == 0 (divide by 1). */
void t1_forster_table(double gamma0, const int32_t dims[4], const double* theta, const double* z, const double* a1,
                      const double* a2, double* rates) {
  size_t k = 0;
  for (int i_th = 0; i_th < dims[0]; ++i_th) {
    const double th = theta[i_th];
    for (int i_z = 0; i_z < dims[1]; ++i_z) {
      const double zsh = z[i_z];
      for (int i_1 = 0; i_1 < dims[2]; ++i_1) {
        const double ash1 = a1[i_1];
        for (int i_2 = 0; i_2 < dims[3]; ++i_2) {
          const double ash2 = a2[i_2];
          const double r1[3] = {ash1, 0, 0};
          const double r2[3] = {ash2 * cos(th), ash2 * sin(th), zsh};
          const double dR[3] = {r1[0] - r2[0], r1[1] - r2[1], r1[2] - r2[2]};
          double       n1 = norm3(r1), n2 = norm3(r2), nd = norm3(dR);
          const double d1 = (n1 > 0) ? n1 : 1.0, d2 = (n2 > 0) ? n2 : 1.0, dd = (nd > 0) ? nd : 1.0;
          const double u1[3] = {r1[0] / d1, r1[1] / d1, r1[2] / d1};
          const double u2[3] = {r2[0] / d2, r2[1] / d2, r2[2] / d2};
          const double ud[3] = {dR[0] / dd, dR[1] / dd, dR[2] / dd};
          const double angle_factor = cos(th) - 3 * dot3(u1, ud) * dot3(u2, ud);
          rates[k++] = gamma0 * (angle_factor * angle_factor) * pow(1.e-9 / nd, 6);
        }
      }
    }
  }
}

/* scatterer.cpp:18-30, literally (unsigned wrap-around of left = -1 included) */
int64_t t1_select(const double* cum, int64_t d, double dice) {
  unsigned left = (unsigned)-1;
  unsigned right = (unsigned)d - 1;
  while (left + 1 < right) {
    unsigned mid = (left + right) / 2;
    if (cum[mid] <= dice) {
      left = mid;
    } else {
      right = mid;
    }
  }
  return (int64_t)right;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* simulation object                                                                                           */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  double  pos[3], orient[3];
  int32_t left, right;
  double  max_rate, inv_max_rate;
} site_t;

typedef struct {
  int32_t site;
  double  pos[3], old_pos[3], delta[3];
  double  ff;
  int32_t heading_right;
  uint64_t gid;   /* global exciton id = Philox stream id / replay list id */
  uint64_t ndraw; /* draws consumed so far */
  int64_t  nevent;
} exciton_t;

typedef struct { int32_t* v; int64_t n, cap; } ivec_t;
static void ivec_push(ivec_t* a, int32_t x) {
  if (a->n == a->cap) {
    a->cap = a->cap ? 2 * a->cap : 16;
    a->v = (int32_t*)realloc(a->v, (size_t)a->cap * sizeof(int32_t));
  }
  a->v[a->n++] = x;
}

struct t1_sim {
  /* table */
  int32_t dims[4];
  double *theta, *z, *a1, *a2, *rates;
  /* sites */
  site_t* sites;
  int64_t N;
  double  lo[3], hi[3];
  double  radius, velocity;
  /* buckets */
  int32_t nb[3];
  int64_t *bucket_start;
  int32_t *bucket_sites;
  int      lazy_rates; /* max_rate of a site is computed when first needed (full-size C4: 5e6 sites) */
  /* memoised rows */
  int      memo;
  int32_t** memo_ids;
  double**  memo_cum;
  int32_t*  memo_len; /* -1: not computed */
  /* injection */
  int32_t* inject;
  int64_t  n_inject;
  double   rem_lo[3], rem_hi[3];
  /* draws */
  int      draw_mode;
  uint64_t seed;
  int64_t  replay_P;
  const int64_t* replay_off;
  const int32_t* replay_flat;
  int      replay_exhausted;
  int      log_draws, trace;
  ivec_t*  draw_log; /* per gid */
  ivec_t*  site_log;
  int64_t  log_n;
  /* excitons */
  exciton_t* ex;
  int64_t    P, P_cap;
  uint64_t   next_gid;
  double     time;
  int64_t    hops, reinjections;
  /* contacts */
  int32_t  n_seg;
  int64_t  c1_pop, c2_pop;
  double*  area;
  int32_t *c1, *c2;
  int64_t  n_c1, n_c2;
};

t1_sim* t1_create(void) { return (t1_sim*)calloc(1, sizeof(t1_sim)); }

static void free_memo(t1_sim* s) {
  if (s->memo_len) {
    for (int64_t i = 0; i < s->N; ++i) {
      free(s->memo_ids[i]);
      free(s->memo_cum[i]);
    }
    free(s->memo_ids); free(s->memo_cum); free(s->memo_len);
    s->memo_ids = NULL; s->memo_cum = NULL; s->memo_len = NULL;
  }
}
static void free_logs(t1_sim* s) {
  for (int64_t i = 0; i < s->log_n; ++i) {
    if (s->draw_log) free(s->draw_log[i].v);
    if (s->site_log) free(s->site_log[i].v);
  }
  free(s->draw_log); free(s->site_log);
  s->draw_log = NULL; s->site_log = NULL; s->log_n = 0;
}
void t1_destroy(t1_sim* s) {
  if (!s) return;
  free_memo(s);
  free_logs(s);
  free(s->theta); free(s->z); free(s->a1); free(s->a2); free(s->rates);
  free(s->sites); free(s->bucket_start); free(s->bucket_sites); free(s->inject); free(s->ex);
  free(s->area); free(s->c1); free(s->c2);
  free(s);
}

static double* dup_d(const double* p, size_t n) {
  double* r = (double*)malloc(n * sizeof(double));
  memcpy(r, p, n * sizeof(double));
  return r;
}

void t1_set_table(t1_sim* s, const int32_t dims[4], const double* theta, const double* z, const double* a1,
                  const double* a2, const double* rates) {
  memcpy(s->dims, dims, sizeof(s->dims));
  s->theta = dup_d(theta, dims[0]);
  s->z = dup_d(z, dims[1]);
  s->a1 = dup_d(a1, dims[2]);
  s->a2 = dup_d(a2, dims[3]);
  s->rates = dup_d(rates, (size_t)dims[0] * dims[1] * dims[2] * dims[3]);
}

/* arma::abs(grid - x).index_min(): first strict minimum from +inf; NaN never wins => 0 */
static int argmin_abs(const double* grid, int n, double x) {
  double best = INFINITY;
  int    idx = 0;
  for (int i = 0; i < n; ++i) {
    const double d = fabs(grid[i] - x);
    if (d < best) {
      best = d;
      idx = i;
    }
  }
  return idx;
}

double t1_get_rate(const t1_sim* s, double theta, double z, double a1, double a2) {
  const int i_th = argmin_abs(s->theta, s->dims[0], theta);
  const int i_z = argmin_abs(s->z, s->dims[1], z);
  const int i_1 = argmin_abs(s->a1, s->dims[2], a1);
  const int i_2 = argmin_abs(s->a2, s->dims[3], a2);
  return s->rates[(((size_t)i_th * s->dims[1] + i_z) * s->dims[2] + i_1) * s->dims[3] + i_2];
}

/* monte_carlo.h:199-271: site n = tube*n_cols + col, positions nm -> m, chain links n-1 / n+1 inside a tube */
void t1_set_mesh(t1_sim* s, int64_t n_tubes, int64_t n_cols, const double* pos_nm, const double* orient) {
  const int64_t N = n_tubes * n_cols;
  free_memo(s);
  free(s->sites);
  s->sites = (site_t*)calloc((size_t)N, sizeof(site_t));
  s->N = N;
  for (int64_t i = 0; i < n_tubes; ++i) {
    for (int64_t j = 0; j < n_cols; ++j) {
      const int64_t n = i * n_cols + j;
      for (int c = 0; c < 3; ++c) {
        s->sites[n].pos[c] = pos_nm[c * N + n] * 1.e-9;
        s->sites[n].orient[c] = orient[c * N + n];
      }
      s->sites[n].left = (j > 0) ? (int32_t)(n - 1) : -1;
      s->sites[n].right = (j + 1 < n_cols) ? (int32_t)(n + 1) : -1;
    }
  }
}

/* monte_carlo.h:730-750 */
static void swap_sites(site_t* s, int i, int j) {
  const int iLeft = s[i].left, iRight = s[i].right, jLeft = s[j].left, jRight = s[j].right;
  const int new_i = j, new_j = i;
  if (iLeft > -1) s[iLeft].right = new_i;
  if (iRight > -1) s[iRight].left = new_i;
  if (jLeft > -1) s[jLeft].right = new_j;
  if (jRight > -1) s[jRight].left = new_j;
  site_t t = s[i];
  s[i] = s[j];
  s[j] = t;
}

/* monte_carlo.h:752-772: scan from the front, swap out-of-box sites with the current tail, cut their links */
void t1_trim(t1_sim* sim, const double xlim[2], const double ylim[2], const double zlim[2]) {
  site_t* s = sim->sites;
  int     j = (int)sim->N;
  for (int i = 0; i < j;) {
    if (s[i].pos[0] < xlim[0] || s[i].pos[1] < ylim[0] || s[i].pos[2] < zlim[0] || s[i].pos[0] > xlim[1] ||
        s[i].pos[1] > ylim[1] || s[i].pos[2] > zlim[1]) {
      --j;
      swap_sites(s, i, j);
      if (s[j].left > -1) s[s[j].left].right = -1;
      if (s[j].right > -1) s[s[j].right].left = -1;
    } else {
      ++i;
    }
  }
  free_memo(sim);
  sim->N = j;
}

/* monte_carlo.h:328-340 */
void t1_find_domain(t1_sim* s) {
  for (int c = 0; c < 3; ++c) s->lo[c] = s->hi[c] = s->sites[0].pos[c];
  for (int64_t i = 0; i < s->N; ++i)
    for (int c = 0; c < 3; ++c) {
      s->lo[c] = s->lo[c] > s->sites[i].pos[c] ? s->sites[i].pos[c] : s->lo[c];
      s->hi[c] = s->hi[c] < s->sites[i].pos[c] ? s->sites[i].pos[c] : s->hi[c];
    }
}

static void cell_of(const t1_sim* s, const double pos[3], int c[3]) {
  for (int k = 0; k < 3; ++k) c[k] = (int)((pos[k] - s->lo[k]) / s->radius);
}

/* monte_carlo.h:375-395: n = ceil(extent/R)+1 cells per axis; bucket contents in site-list order */
void t1_build_buckets(t1_sim* s, double radius) {
  s->radius = radius;
  for (int k = 0; k < 3; ++k) s->nb[k] = (int)(ceil((s->hi[k] - s->lo[k]) / radius) + 1);
  const int64_t nc = (int64_t)s->nb[0] * s->nb[1] * s->nb[2];
  free(s->bucket_start);
  free(s->bucket_sites);
  s->bucket_start = (int64_t*)calloc((size_t)nc + 1, sizeof(int64_t));
  s->bucket_sites = (int32_t*)malloc((size_t)s->N * sizeof(int32_t));
  for (int64_t i = 0; i < s->N; ++i) {
    int c[3];
    cell_of(s, s->sites[i].pos, c);
    s->bucket_start[(int64_t)c[0] + (int64_t)c[1] * s->nb[0] + (int64_t)c[2] * s->nb[0] * s->nb[1] + 1]++;
  }
  for (int64_t b = 0; b < nc; ++b) s->bucket_start[b + 1] += s->bucket_start[b];
  int64_t* fill = (int64_t*)malloc((size_t)nc * sizeof(int64_t));
  memcpy(fill, s->bucket_start, (size_t)nc * sizeof(int64_t));
  for (int64_t i = 0; i < s->N; ++i) {
    int c[3];
    cell_of(s, s->sites[i].pos, c);
    s->bucket_sites[fill[(int64_t)c[0] + (int64_t)c[1] * s->nb[0] + (int64_t)c[2] * s->nb[0] * s->nb[1]]++] = (int32_t)i;
  }
  free(fill);
  free_memo(s);
}

void t1_set_velocity(t1_sim* s, double v) { s->velocity = v; }

/* scatterer.cpp:34-83.  Candidates: stencil cells in the order x outer, y, z inner (monte_carlo.h:402-411), sites of
 * a cell in list order.  Returns the row length; ids/cum hold the first min(d, cap) entries. */
static int64_t find_neighbors(const t1_sim* s, int64_t i, int32_t* ids, double* cum, int64_t cap) {
  const site_t* s1 = &s->sites[i];
  int           c[3];
  cell_of(s, s1->pos, c);
  int64_t d = 0;
  double  acc = 0;
  for (int ix = c[0] - 1; ix <= c[0] + 1; ++ix)
    for (int iy = c[1] - 1; iy <= c[1] + 1; ++iy)
      for (int iz = c[2] - 1; iz <= c[2] + 1; ++iz) {
        if (!(ix > -1 && ix < s->nb[0] && iy > -1 && iy < s->nb[1] && iz > -1 && iz < s->nb[2])) continue;
        const int64_t b = (int64_t)ix + (int64_t)iy * s->nb[0] + (int64_t)iz * s->nb[0] * s->nb[1];
        for (int64_t q = s->bucket_start[b]; q < s->bucket_start[b + 1]; ++q) {
          const int32_t j = s->bucket_sites[q];
          const site_t* s2 = &s->sites[j];
          const double  dR[3] = {s1->pos[0] - s2->pos[0], s1->pos[1] - s2->pos[1], s1->pos[2] - s2->pos[2]};
          const double  distance = norm3(dR);
          if ((distance < s->radius) && (distance > 0.4e-9)) {
            const double* a1 = s1->orient;
            const double* a2 = s2->orient;
            const double  cosTheta = dot3(a1, a2);
            double        theta, axis_shift_1, axis_shift_2, z_shift;
            if (cosTheta == 1) {
              axis_shift_1 = 0;
              axis_shift_2 = dot3(dR, a1);
              theta = 0;
              const double t = dot3(dR, a1);
              const double v[3] = {dR[0] - a1[0] * t, dR[1] - a1[1] * t, dR[2] - a1[2] * t};
              z_shift = norm3(v);
            } else {
              theta = acos(cosTheta);
              const double y1 = dot3(a1, dR);
              const double y2 = dot3(a2, dR);
              const double sin2Theta = 1 - cosTheta * cosTheta;
              axis_shift_1 = (y1 + y2 * cosTheta) / sin2Theta;
              axis_shift_2 = (y2 + y1 * cosTheta) / sin2Theta;
              const double v[3] = {(a1[0] * axis_shift_1 + s1->pos[0]) - (a2[0] * axis_shift_2 + s2->pos[0]),
                                   (a1[1] * axis_shift_1 + s1->pos[1]) - (a2[1] * axis_shift_2 + s2->pos[1]),
                                   (a1[2] * axis_shift_1 + s1->pos[2]) - (a2[2] * axis_shift_2 + s2->pos[2])};
              z_shift = norm3(v);
            }
            const double rate = t1_get_rate(s, theta, z_shift, axis_shift_1, axis_shift_2);
            acc = (d == 0) ? rate : acc + rate; /* scatterer.cpp:78-80: sequential prefix sum */
            if (d < cap) {
              ids[d] = j;
              cum[d] = acc;
            }
            ++d;
          }
        }
      }
  return d;
}

/* row access with optional memoisation (the row is a pure function of the site) */
static int64_t get_row(t1_sim* s, int64_t i, const int32_t** ids, const double** cum, int32_t* tmp_ids, double* tmp_cum,
                       int64_t cap) {
  if (!s->memo) {
    const int64_t d = find_neighbors(s, i, tmp_ids, tmp_cum, cap);
    if (d > cap) {
      fprintf(stderr, "[oracle/T1] row longer than scratch (%lld > %lld)\n", (long long)d, (long long)cap);
      abort();
    }
    *ids = tmp_ids;
    *cum = tmp_cum;
    return d;
  }
  if (!s->memo_len) {
    s->memo_len = (int32_t*)malloc((size_t)s->N * sizeof(int32_t));
    s->memo_ids = (int32_t**)calloc((size_t)s->N, sizeof(int32_t*));
    s->memo_cum = (double**)calloc((size_t)s->N, sizeof(double*));
    for (int64_t k = 0; k < s->N; ++k) s->memo_len[k] = -1;
  }
  if (s->memo_len[i] < 0) {
    const int64_t d = find_neighbors(s, i, tmp_ids, tmp_cum, cap);
    if (d > cap) abort();
    s->memo_ids[i] = (int32_t*)malloc((size_t)(d ? d : 1) * sizeof(int32_t));
    s->memo_cum[i] = (double*)malloc((size_t)(d ? d : 1) * sizeof(double));
    memcpy(s->memo_ids[i], tmp_ids, (size_t)d * sizeof(int32_t));
    memcpy(s->memo_cum[i], tmp_cum, (size_t)d * sizeof(double));
    s->memo_len[i] = (int32_t)d;
  }
  *ids = s->memo_ids[i];
  *cum = s->memo_cum[i];
  return s->memo_len[i];
}

#define ROW_CAP 8192

void t1_set_memo(t1_sim* s, int on) {
  s->memo = on;
  if (!on) free_memo(s);
}

/* scatterer.h:89-93 over all sites (monte_carlo.h:426-440); the reference is undefined for an empty row */
void t1_set_lazy_rates(t1_sim* s, int on) { s->lazy_rates = on; }
/* the rate fields of one site, exactly as the loop below computes them */
static void ensure_rate(t1_sim* s, int64_t i) {
  if (s->sites[i].inv_max_rate == s->sites[i].inv_max_rate) return; /* not NaN: known */
  int32_t* ids = (int32_t*)malloc(ROW_CAP * sizeof(int32_t));
  double*  cum = (double*)malloc(ROW_CAP * sizeof(double));
  const int64_t d = find_neighbors(s, i, ids, cum, ROW_CAP);
  if (d == 0 || d > ROW_CAP) {
    fprintf(stderr, "[oracle/T1] a site has no neighbour (undefined behaviour in scatterer.h:91) or a row > %d\n", ROW_CAP);
    abort();
  }
  s->sites[i].max_rate = cum[d - 1];
  s->sites[i].inv_max_rate = 1. / s->sites[i].max_rate;
  free(ids);
  free(cum);
}
void t1_set_max_rate(t1_sim* s) {
  int bad = 0;
  if (s->lazy_rates) {
    for (int64_t i = 0; i < s->N; ++i) s->sites[i].max_rate = s->sites[i].inv_max_rate = NAN;
    return;
  }
#pragma omp parallel
  {
    int32_t* ids = (int32_t*)malloc(ROW_CAP * sizeof(int32_t));
    double*  cum = (double*)malloc(ROW_CAP * sizeof(double));
#pragma omp for schedule(dynamic, 64)
    for (int64_t i = 0; i < s->N; ++i) {
      const int64_t d = find_neighbors(s, i, ids, cum, ROW_CAP);
      if (d == 0 || d > ROW_CAP) {
        bad = 1;
        continue;
      }
      s->sites[i].max_rate = cum[d - 1];
      s->sites[i].inv_max_rate = 1. / s->sites[i].max_rate;
    }
    free(ids);
    free(cum);
  }
  if (bad) {
    fprintf(stderr, "[oracle/T1] a site has no neighbour (undefined behaviour in scatterer.h:91) or a row > %d\n", ROW_CAP);
    abort();
  }
}

/* monte_carlo.cpp:203-251 */
void t1_injection(t1_sim* s, int32_t n) {
  double x[64], y[64], z[64];
  if (n < 1 || n > 63 || n % 2 != 1) abort();
  const double dx = (s->hi[0] - s->lo[0]) / (double)n, dy = (s->hi[1] - s->lo[1]) / (double)n,
               dz = (s->hi[2] - s->lo[2]) / (double)n;
  for (int i = 0; i <= n; ++i) {
    x[i] = (double)i * dx + s->lo[0];
    y[i] = (double)i * dy + s->lo[1];
    z[i] = (double)i * dz + s->lo[2];
  }
  free(s->inject);
  s->inject = (int32_t*)malloc((size_t)s->N * sizeof(int32_t));
  s->n_inject = 0;
  for (int64_t i = 0; i < s->N; ++i) {
    const double* p = s->sites[i].pos;
    if (x[n / 2] <= p[0] && p[0] <= x[n / 2 + 1] && y[n / 2] <= p[1] && p[1] <= y[n / 2 + 1] && z[n / 2] <= p[2] &&
        p[2] <= z[n / 2 + 1])
      s->inject[s->n_inject++] = (int32_t)i;
  }
  s->rem_lo[0] = x[1]; s->rem_lo[1] = y[1]; s->rem_lo[2] = z[1];
  s->rem_hi[0] = x[n - 1]; s->rem_hi[1] = y[n - 1]; s->rem_hi[2] = z[n - 1];
}

int64_t t1_num_sites(const t1_sim* s) { return s->N; }
void    t1_sites(const t1_sim* s, double* pos, double* orient, int32_t* left, int32_t* right, double* max_rate,
                 double* inv_max_rate) {
  for (int64_t i = 0; i < s->N; ++i) {
    for (int c = 0; c < 3; ++c) {
      pos[c * s->N + i] = s->sites[i].pos[c];
      orient[c * s->N + i] = s->sites[i].orient[c];
    }
    left[i] = s->sites[i].left;
    right[i] = s->sites[i].right;
    max_rate[i] = s->sites[i].max_rate;
    inv_max_rate[i] = s->sites[i].inv_max_rate;
  }
}
void t1_domain(const t1_sim* s, double d[6]) {
  for (int c = 0; c < 3; ++c) { d[c] = s->lo[c]; d[3 + c] = s->hi[c]; }
}
void t1_removal_domain(const t1_sim* s, double d[6]) {
  for (int c = 0; c < 3; ++c) { d[c] = s->rem_lo[c]; d[3 + c] = s->rem_hi[c]; }
}
int64_t t1_num_inject(const t1_sim* s) { return s->n_inject; }
void    t1_inject(const t1_sim* s, int32_t* ids) { memcpy(ids, s->inject, (size_t)s->n_inject * sizeof(int32_t)); }
void    t1_bucket_dims(const t1_sim* s, int32_t n[3]) { memcpy(n, s->nb, sizeof(s->nb)); }
int64_t t1_row(const t1_sim* s, int64_t i, int32_t* ids, double* cum, int64_t cap) { return find_neighbors(s, i, ids, cum, cap); }

void t1_degrees(const t1_sim* s, int32_t* deg) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < s->N; ++i) deg[i] = (int32_t)find_neighbors(s, i, NULL, NULL, 0);
}
void t1_csr(const t1_sim* s, const int64_t* row_ptr, int32_t* ids, double* cum) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < s->N; ++i)
    find_neighbors(s, i, ids + row_ptr[i], cum + row_ptr[i], row_ptr[i + 1] - row_ptr[i]);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* draws                                                                                                       */
/* ------------------------------------------------------------------------------------------------------------ */
void t1_draws_glibc(t1_sim* s) { s->draw_mode = T1_DRAWS_GLIBC; }
void t1_draws_philox(t1_sim* s, uint64_t seed) {
  s->draw_mode = T1_DRAWS_PHILOX;
  s->seed = seed;
}
void t1_draws_replay(t1_sim* s, int64_t P, const int64_t* offsets, const int32_t* flat) {
  s->draw_mode = T1_DRAWS_REPLAY;
  s->replay_P = P;
  s->replay_off = offsets;
  s->replay_flat = flat;
  s->replay_exhausted = 0;
}
int  t1_replay_exhausted(const t1_sim* s) { return s->replay_exhausted; }
void t1_log_draws(t1_sim* s, int on) { s->log_draws = on; }
void t1_trace_sites(t1_sim* s, int on) { s->trace = on; }

static void ensure_logs(t1_sim* s, uint64_t gid) {
  if ((int64_t)gid < s->log_n) return;
  int64_t n = s->log_n ? s->log_n : 64;
  while (n <= (int64_t)gid) n *= 2;
  s->draw_log = (ivec_t*)realloc(s->draw_log, (size_t)n * sizeof(ivec_t));
  s->site_log = (ivec_t*)realloc(s->site_log, (size_t)n * sizeof(ivec_t));
  memset(s->draw_log + s->log_n, 0, (size_t)(n - s->log_n) * sizeof(ivec_t));
  memset(s->site_log + s->log_n, 0, (size_t)(n - s->log_n) * sizeof(ivec_t));
  s->log_n = n;
}

/* one 31-bit draw for exciton e (the reference calls rand()) */
static int32_t draw(t1_sim* s, exciton_t* e) {
  int32_t r;
  switch (s->draw_mode) {
    case T1_DRAWS_PHILOX:
      r = t1_philox_draw(s->seed, e->gid, e->ndraw);
      break;
    case T1_DRAWS_REPLAY: {
      const int64_t at = s->replay_off[e->gid] + (int64_t)e->ndraw;
      if ((int64_t)e->gid >= s->replay_P || at >= s->replay_off[e->gid + 1]) {
        s->replay_exhausted = 1;
        r = 1;
      } else {
        r = s->replay_flat[at];
      }
      break;
    }
    default:
      r = (int32_t)rand();
  }
  e->ndraw++;
  if (s->log_draws) {
    ensure_logs(s, e->gid);
    ivec_push(&s->draw_log[e->gid], r);
  }
  return r;
}

/* scatterer.h:74-80 */
static double ff_time(t1_sim* s, exciton_t* e, int32_t site) {
  int32_t r;
  while ((r = draw(s, e)) == 0) {
  }
  if (s->lazy_rates) ensure_rate(s, site);
  return -s->sites[site].inv_max_rate * log((double)r / (double)T1_RAND_MAX);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* exciton kinetics                                                                                            */
/* ------------------------------------------------------------------------------------------------------------ */
/* particle.cpp:9-54 */
static void fly(const t1_sim* s, exciton_t* e, double dt) {
  const site_t* sites = s->sites;
  if (sites[e->site].left < 0 && sites[e->site].right < 0) return;
  for (;;) {
    const site_t* cur = &sites[e->site];
    int           next;
    if (e->heading_right) {
      next = (cur->right > -1) ? cur->right : cur->left;
    } else {
      next = (cur->left > -1) ? cur->left : cur->right;
    }
    e->heading_right = (next == cur->right) ? 1 : 0;
    const double* t = sites[next].pos;
    const double  dv[3] = {e->pos[0] - t[0], e->pos[1] - t[1], e->pos[2] - t[2]};
    const double  dist = norm3(dv);
    if (dist / s->velocity < dt) {
      e->pos[0] = t[0]; e->pos[1] = t[1]; e->pos[2] = t[2];
      e->site = next;
      dt -= dist / s->velocity;
    } else {
      const double w[3] = {t[0] - e->pos[0], t[1] - e->pos[1], t[2] - e->pos[2]};
      const double n = norm3(w);
      const double nn = (n > 0) ? n : 1.0;
      const double k = s->velocity * dt;
      for (int c = 0; c < 3; ++c) e->pos[c] += (w[c] / nn) * k;
      return;
    }
  }
}

/* scatterer.cpp:9-31 */
static int32_t update_state(t1_sim* s, exciton_t* e, int32_t* tmp_ids, double* tmp_cum) {
  const int32_t* ids;
  const double*  cum;
  const int64_t  d = get_row(s, e->site, &ids, &cum, tmp_ids, tmp_cum, ROW_CAP);
  if (d == 0) return e->site;
  const double dice = cum[d - 1] * (double)draw(s, e) / (double)T1_RAND_MAX;
  s->hops++;
  e->nevent++;
  return ids[t1_select(cum, d, dice)];
}

/* particle.cpp:57-80 */
static void particle_step(t1_sim* s, exciton_t* e, double dt, int32_t* tmp_ids, double* tmp_cum) {
  for (int c = 0; c < 3; ++c) e->old_pos[c] = e->pos[c];
  while (e->ff <= dt) {
    dt -= e->ff;
    fly(s, e, e->ff);
    const int32_t n = update_state(s, e, tmp_ids, tmp_cum);
    if (n != e->site) {
      e->site = n;
      for (int c = 0; c < 3; ++c) e->pos[c] = s->sites[n].pos[c];
    }
    if (s->trace) {
      ensure_logs(s, e->gid);
      ivec_push(&s->site_log[e->gid], e->site);
    }
    e->ff = ff_time(s, e, e->site);
  }
  fly(s, e, dt);
  e->ff -= dt;
}

static exciton_t* push_exciton(t1_sim* s) {
  if (s->P == s->P_cap) {
    s->P_cap = s->P_cap ? 2 * s->P_cap : 1024;
    s->ex = (exciton_t*)realloc(s->ex, (size_t)s->P_cap * sizeof(exciton_t));
  }
  return &s->ex[s->P++];
}

/* particle ctor, particle.h:48-52 (the site draw precedes it at every call site) */
static void make_exciton(t1_sim* s, exciton_t* e, uint64_t gid, const int32_t* site_list, int64_t n_list) {
  memset(e, 0, sizeof(*e));
  e->gid = gid;
  const int32_t dice = draw(s, e) % (int32_t)n_list;
  e->site = site_list[dice];
  for (int c = 0; c < 3; ++c) e->pos[c] = e->old_pos[c] = s->sites[e->site].pos[c];
  e->ff = ff_time(s, e, e->site);
  e->heading_right = draw(s, e) % 2;
}

/* monte_carlo.cpp:308-316 */
void t1_kubo_create_particles(t1_sim* s, int64_t P, uint64_t first_global_id) {
  s->P = 0;
  s->time = 0;
  s->hops = s->reinjections = 0;
  free_logs(s);
  for (int64_t i = 0; i < P; ++i) make_exciton(s, push_exciton(s), first_global_id + (uint64_t)i, s->inject, s->n_inject);
  s->next_gid = first_global_id + (uint64_t)P;
}

/* monte_carlo.cpp:319-342 followed by the ensemble average of :396-406 */
void t1_kubo_step(t1_sim* s, double dt, int64_t nsteps, double* msd) {
  int32_t* tmp_ids = (int32_t*)malloc(ROW_CAP * sizeof(int32_t));
  double*  tmp_cum = (double*)malloc(ROW_CAP * sizeof(double));
  for (int64_t st = 0; st < nsteps; ++st) {
    for (int64_t i = 0; i < s->P; ++i) {
      exciton_t* e = &s->ex[i];
      particle_step(s, e, dt, tmp_ids, tmp_cum);
      for (int c = 0; c < 3; ++c) e->delta[c] += e->pos[c] - e->old_pos[c];
      if (e->pos[0] < s->rem_lo[0] || e->pos[1] < s->rem_lo[1] || e->pos[2] < s->rem_lo[2] || s->rem_hi[0] < e->pos[0] ||
          s->rem_hi[1] < e->pos[1] || s->rem_hi[2] < e->pos[2]) {
        const int32_t dice = draw(s, e) % (int32_t)s->n_inject;
        e->site = s->inject[dice];
        for (int c = 0; c < 3; ++c) e->pos[c] = s->sites[e->site].pos[c];
        s->reinjections++;
      }
    }
    s->time += dt;
    if (msd) {
      double a[3] = {0, 0, 0};
      for (int64_t i = 0; i < s->P; ++i)
        for (int c = 0; c < 3; ++c) a[c] += s->ex[i].delta[c] * s->ex[i].delta[c];
      for (int c = 0; c < 3; ++c) msd[3 * st + c] = a[c] / (double)s->P;
    }
  }
  free(tmp_ids);
  free(tmp_cum);
}

int64_t t1_num_particles(const t1_sim* s) { return s->P; }
void    t1_particles(const t1_sim* s, int32_t* site, double* pos, double* old_pos, double* delta, double* ff,
                     int32_t* heading) {
  for (int64_t i = 0; i < s->P; ++i) {
    site[i] = s->ex[i].site;
    for (int c = 0; c < 3; ++c) {
      pos[c * s->P + i] = s->ex[i].pos[c];
      if (old_pos) old_pos[c * s->P + i] = s->ex[i].old_pos[c];
      delta[c * s->P + i] = s->ex[i].delta[c];
    }
    ff[i] = s->ex[i].ff;
    heading[i] = s->ex[i].heading_right;
  }
}
double  t1_time(const t1_sim* s) { return s->time; }
int64_t t1_hops(const t1_sim* s) { return s->hops; }
int64_t t1_reinjections(const t1_sim* s) { return s->reinjections; }
void    t1_event_counts(const t1_sim* s, int64_t* out) {
  for (int64_t i = 0; i < s->P; ++i) out[i] = s->ex[i].nevent;
}
/* logs are indexed by gid - (smallest gid ever created is assumed 0-based by the caller) */
void t1_draw_counts(const t1_sim* s, int64_t* out) {
  for (int64_t i = 0; i < (int64_t)s->next_gid; ++i) out[i] = (i < s->log_n && s->draw_log) ? s->draw_log[i].n : 0;
}
void t1_logged_draws(const t1_sim* s, int32_t* flat) {
  int64_t k = 0;
  for (int64_t i = 0; i < (int64_t)s->next_gid && i < s->log_n; ++i)
    for (int64_t q = 0; q < s->draw_log[i].n; ++q) flat[k++] = s->draw_log[i].v[q];
}
void t1_trace_counts(const t1_sim* s, int64_t* out) {
  for (int64_t i = 0; i < (int64_t)s->next_gid; ++i) out[i] = (i < s->log_n && s->site_log) ? s->site_log[i].n : 0;
}
void t1_traced_sites(const t1_sim* s, int32_t* flat) {
  int64_t k = 0;
  for (int64_t i = 0; i < (int64_t)s->next_gid && i < s->log_n; ++i)
    for (int64_t q = 0; q < s->site_log[i].n; ++q) flat[k++] = s->site_log[i].v[q];
}

/* ------------------------------------------------------------------------------------------------------------ */
/* contact flavour                                                                                             */
/* ------------------------------------------------------------------------------------------------------------ */
/* monte_carlo.h:646-688 (note the if / else-if: a site updates the min OR the max of its slab, never both) */
static void get_area(t1_sim* s) {
  const int n_seg = s->n_seg;
  const double ymax = s->hi[1], ymin = s->lo[1];
  const double dy = (ymax - ymin) / (double)n_seg;
  double *xmax = (double*)malloc(4 * (size_t)n_seg * sizeof(double)), *xmin = xmax + n_seg, *zmax = xmin + n_seg,
         *zmin = zmax + n_seg;
  for (int i = 0; i < n_seg; ++i) {
    xmax[i] = s->lo[0];
    xmin[i] = s->hi[0];
    zmax[i] = s->lo[2];
    zmin[i] = s->hi[2];
  }
  for (int64_t k = 0; k < s->N; ++k) {
    const double* p = s->sites[k].pos;
    int           i = (int)((p[1] - ymin) / dy);
    i = i < 0 ? 0 : (i < n_seg ? i : n_seg - 1);
    if (xmin[i] > p[0]) {
      xmin[i] = p[0];
    } else if (xmax[i] < p[0]) {
      xmax[i] = p[0];
    }
    if (zmin[i] > p[2]) {
      zmin[i] = p[2];
    } else if (zmax[i] < p[2]) {
      zmax[i] = p[2];
    }
  }
  free(s->area);
  s->area = (double*)malloc((size_t)n_seg * sizeof(double));
  for (int i = 0; i < n_seg; ++i) s->area[i] = (zmax[i] - zmin[i]) * (xmax[i] - xmin[i]);
  free(xmax);
}

/* monte_carlo.h:494-516 */
static int64_t contact_scats(const t1_sim* s, int i, int32_t** out) {
  const double ymin = s->lo[1], ymax = s->hi[1];
  const double dy = (ymax - ymin) / (double)s->n_seg;
  const double y1 = ymin + (double)(i - 1) * dy;
  const double y2 = ymin + (double)i * dy;
  int32_t*     l = (int32_t*)malloc((size_t)(s->N ? s->N : 1) * sizeof(int32_t));
  int64_t      n = 0;
  for (int64_t k = 0; k < s->N; ++k)
    if (s->sites[k].pos[1] >= y1 && s->sites[k].pos[1] <= y2) l[n++] = (int32_t)k;
  *out = l;
  return n;
}

/* monte_carlo.h:157-195 after the common set-up, then create_particles :274-316.
 * Call order expected from the host: set_table, set_mesh, trim, find_domain, [contacts_init does area], buckets,
 * set_max_rate must precede this (ff_time needs the rates) -- the reference computes area before the buckets but the
 * two are independent. */
void t1_contacts_init(t1_sim* s, int32_t n_seg, int64_t c1_pop, int64_t c2_pop) {
  s->n_seg = n_seg;
  s->c1_pop = c1_pop;
  s->c2_pop = c2_pop;
  get_area(s);
  free(s->c1);
  free(s->c2);
  s->n_c1 = contact_scats(s, 1, &s->c1);
  s->n_c2 = contact_scats(s, n_seg, &s->c2);
  /* create_particles(domain, n_seg, scat_list, left_pop = c1_pop, right_pop = c2_pop) */
  s->P = 0;
  s->time = 0;
  s->hops = s->reinjections = 0;
  s->next_gid = 0;
  free_logs(s);
  const double y_min = s->lo[1], y_max = s->hi[1];
  const double dy = (y_max - y_min) / (double)n_seg;
  const double dp = (double)(c2_pop - c1_pop) / ((double)n_seg - 1);
  int32_t*     s_list = (int32_t*)malloc((size_t)(s->N ? s->N : 1) * sizeof(int32_t));
  for (int i = 0; i < n_seg; ++i) {
    const int    n_particle = (int)round((double)c1_pop + (double)i * dp);
    const double y1 = y_min + (double)i * dy;
    const double y2 = y1 + dy;
    int64_t      n_list = 0;
    for (int64_t k = 0; k < s->N; ++k)
      if (y1 <= s->sites[k].pos[1] && s->sites[k].pos[1] < y2) s_list[n_list++] = (int32_t)k;
    for (int n = 0; n < n_particle; ++n) make_exciton(s, push_exciton(s), s->next_gid++, s_list, n_list);
  }
  free(s_list);
}

void    t1_area(const t1_sim* s, double* area) { memcpy(area, s->area, (size_t)s->n_seg * sizeof(double)); }
int64_t t1_num_contact_sites(const t1_sim* s, int which) { return which == 1 ? s->n_c1 : s->n_c2; }
void    t1_contact_sites(const t1_sim* s, int which, int32_t* ids) {
  memcpy(ids, which == 1 ? s->c1 : s->c2, (size_t)(which == 1 ? s->n_c1 : s->n_c2) * sizeof(int32_t));
}

/* monte_carlo.h:458-491 */
static void repopulate(t1_sim* s, double ymin, double ymax, int64_t n_particle, const int32_t* s_list, int64_t n_list) {
  int64_t j = s->P;
  for (int64_t i = 0; i < j;) {
    if (s->ex[i].pos[1] >= ymin && s->ex[i].pos[1] <= ymax) {
      --j;
      exciton_t t = s->ex[i];
      s->ex[i] = s->ex[j];
      s->ex[j] = t;
    } else {
      ++i;
    }
  }
  int64_t       n = 0;
  const int64_t final_size = j + n_particle;
  const int64_t j_lim = s->P < final_size ? s->P : final_size;
  for (; j < j_lim; ++j) {
    make_exciton(s, &s->ex[j], s->next_gid++, s_list, n_list);
    ++n;
  }
  for (; n < n_particle; ++n) make_exciton(s, push_exciton(s), s->next_gid++, s_list, n_list);
  s->P = final_size;
}

/* main.cpp:98-106: step(dt); save_metrics(dt); repopulate_contacts() -- the metrics are returned as raw counts
 * (population per slab, net crossings per interface) instead of being divided by area and written to file. */
void t1_contact_iteration(t1_sim* s, double dt, int64_t* pop, int64_t* curr) {
  int32_t* tmp_ids = (int32_t*)malloc(ROW_CAP * sizeof(int32_t));
  double*  tmp_cum = (double*)malloc(ROW_CAP * sizeof(double));
  const int n = s->n_seg;
  /* step: monte_carlo.h:343-355 */
  for (int64_t i = 0; i < s->P; ++i) particle_step(s, &s->ex[i], dt, tmp_ids, tmp_cum);
  s->time += dt;
  free(tmp_ids);
  free(tmp_cum);
  const double ymax = s->hi[1], ymin = s->lo[1];
  const double dy = (ymax - ymin) / (double)n;
  /* save_population_profile: monte_carlo.h:566-573 */
  if (pop) {
    for (int i = 0; i < n; ++i) pop[i] = 0;
    for (int64_t k = 0; k < s->P; ++k) {
      int i = (int)((s->ex[k].pos[1] - ymin) / dy);
      i = i < 0 ? 0 : (i < n ? i : n - 1);
      pop[i]++;
    }
  }
  /* save_currents: monte_carlo.h:593-636 */
  if (curr) {
    for (int i = 1; i < n; ++i) {
      const double y = ymin + dy * (double)i;
      int64_t      c = 0;
      for (int64_t k = 0; k < s->P; ++k) {
        if (s->ex[k].old_pos[1] < y && s->ex[k].pos[1] >= y) {
          c++;
        } else if (s->ex[k].old_pos[1] >= y && s->ex[k].pos[1] < y) {
          c--;
        }
      }
      curr[i - 1] = c;
    }
  }
  /* repopulate_contacts: monte_carlo.h:443-455 */
  repopulate(s, ymin, ymin + dy, s->c1_pop, s->c1, s->n_c1);
  repopulate(s, ymin + (double)(n - 1) * dy, ymax, s->c2_pop, s->c2, s->n_c2);
}

/* monte_carlo::track_particle (monte_carlo.h:786-818): one exciton created on the first contact by repopulate(...,1,...),
 * stepped while pos.y < ymin + (n_seg-1)*dy; path receives the position after every step.  max_steps bounds the
 * reference's unbounded loop.  Returns the number of rows; *reached tells whether the last slab was entered. */
int64_t t1_track_particle(t1_sim* s, double dt, uint64_t gid, int64_t max_steps, double* path, int32_t* reached) {
  int32_t*     tmp_ids = (int32_t*)malloc(ROW_CAP * sizeof(int32_t));
  double*      tmp_cum = (double*)malloc(ROW_CAP * sizeof(double));
  const double ymin = s->lo[1], ymax = s->hi[1];
  const double dy = (ymax - ymin) / (double)s->n_seg;
  const double y1 = ymin + (double)(s->n_seg - 1) * dy;
  exciton_t    e;
  make_exciton(s, &e, gid, s->c1, s->n_c1);
  int64_t n = 0;
  while (e.pos[1] < y1 && n < max_steps) {
    particle_step(s, &e, dt, tmp_ids, tmp_cum);
    for (int c = 0; c < 3; ++c) path[3 * n + c] = e.pos[c];
    ++n;
  }
  if (reached) *reached = !(e.pos[1] < y1);
  free(tmp_ids);
  free(tmp_cum);
  return n;
}

/* log(r / RAND_MAX) with the host libm, for every recorded draw: fed to the engine's replay mode so that free-flight
 * times are bit-identical to a glibc run (scatterer.h:79). */
void t1_log_ratios(const int32_t* draws, int64_t n, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = log((double)draws[i] / (double)T1_RAND_MAX);
}
