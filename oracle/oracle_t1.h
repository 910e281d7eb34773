/* oracle/oracle_t1.h -- TEST INFRASTRUCTURE ONLY: the "T1" CPU oracle.
 *
 * A plain-C restatement of the reference's exciton hop path (see oracle_t1.c for the file:line map).  It exists to
 * check the CUDA engine; the product never links, loads or calls it.  Parity pinning: T1 with glibc draws is proven
 * bit-identical to the reference's own code (oracle/_ref/libt0.so, built from /root/reference) by
 * tests/test_oracle_t0.py, and against the golden vectors under tests/golden/ which that library generated.
 */
#ifndef CNTMC_ORACLE_T1_H
#define CNTMC_ORACLE_T1_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct t1_sim t1_sim;

enum { T1_DRAWS_GLIBC = 0, T1_DRAWS_PHILOX = 1, T1_DRAWS_REPLAY = 2 };

/* ---- stand-alone pieces (known-answer testable) ------------------------------------------------------------ */
void   t1_linspace(double start, double end, int64_t n, double* out);
void   t1_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void   t1_philox2x32_10(const uint32_t ctr[2], uint32_t key, uint32_t out[2]);
int32_t t1_philox_draw(uint64_t seed, uint64_t exciton, uint64_t k); /* 31-bit draw k of an exciton's stream */
void   t1_forster_table(double gamma0, const int32_t dims[4], const double* theta, const double* z, const double* a1,
                        const double* a2, double* rates);
int64_t t1_select(const double* cum, int64_t d, double dice); /* scatterer.cpp:18-30 */
void   t1_log_ratios(const int32_t* draws, int64_t n, double* out); /* host libm log(r/RAND_MAX) */

/* ---- simulation object ------------------------------------------------------------------------------------- */
t1_sim* t1_create(void);
void    t1_destroy(t1_sim* s);

void t1_set_table(t1_sim* s, const int32_t dims[4], const double* theta, const double* z, const double* a1,
                  const double* a2, const double* rates /* [theta][z][a1][a2] */);
double t1_get_rate(const t1_sim* s, double theta, double z, double a1, double a2);

/* create_scatterers: pos_nm / orient are [3][n_tubes*n_cols], tube-major */
void    t1_set_mesh(t1_sim* s, int64_t n_tubes, int64_t n_cols, const double* pos_nm, const double* orient);
void    t1_trim(t1_sim* s, const double xlim[2], const double ylim[2], const double zlim[2]);
void    t1_find_domain(t1_sim* s);
void    t1_build_buckets(t1_sim* s, double radius);
void    t1_set_velocity(t1_sim* s, double v);
void    t1_set_lazy_rates(t1_sim* s, int on); /* before t1_set_max_rate: compute a site's rate when first needed (same value) */
void    t1_set_max_rate(t1_sim* s); /* one find_neighbors per site; aborts on an empty row like the reference's UB */
void    t1_injection(t1_sim* s, int32_t n_sections);
int64_t t1_num_sites(const t1_sim* s);
void    t1_sites(const t1_sim* s, double* pos, double* orient, int32_t* left, int32_t* right, double* max_rate,
                 double* inv_max_rate);
void    t1_domain(const t1_sim* s, double lo_hi[6]);
void    t1_removal_domain(const t1_sim* s, double lo_hi[6]);
int64_t t1_num_inject(const t1_sim* s);
void    t1_inject(const t1_sim* s, int32_t* ids);
void    t1_bucket_dims(const t1_sim* s, int32_t n[3]);
int64_t t1_row(const t1_sim* s, int64_t i, int32_t* ids, double* cum, int64_t cap);
void    t1_degrees(const t1_sim* s, int32_t* deg);
void    t1_csr(const t1_sim* s, const int64_t* row_ptr, int32_t* ids, double* cum);
void    t1_set_memo(t1_sim* s, int on); /* cache rows (pure function of the site) instead of rebuilding per hop */

/* draws */
void t1_draws_glibc(t1_sim* s);                  /* global sequential stream: rand() */
void t1_draws_philox(t1_sim* s, uint64_t seed);  /* per-exciton counter-based streams */
void t1_draws_replay(t1_sim* s, int64_t P, const int64_t* offsets, const int32_t* flat);
void t1_log_draws(t1_sim* s, int on);
void t1_trace_sites(t1_sim* s, int on); /* record the destination site of every scattering event */

/* excitons (Green-Kubo flavour) */
void    t1_kubo_create_particles(t1_sim* s, int64_t P, uint64_t first_global_id);
void    t1_kubo_step(t1_sim* s, double dt, int64_t nsteps, double* msd /* [nsteps][3] or NULL */);
int64_t t1_num_particles(const t1_sim* s);
void    t1_particles(const t1_sim* s, int32_t* site, double* pos, double* old_pos, double* delta, double* ff,
                     int32_t* heading);
double  t1_time(const t1_sim* s);
int64_t t1_hops(const t1_sim* s);       /* scattering events with a non-empty row */
int64_t t1_reinjections(const t1_sim* s);
void    t1_event_counts(const t1_sim* s, int64_t* per_exciton);
void    t1_draw_counts(const t1_sim* s, int64_t* per_exciton);
void    t1_logged_draws(const t1_sim* s, int32_t* flat);
void    t1_trace_counts(const t1_sim* s, int64_t* per_exciton);
void    t1_traced_sites(const t1_sim* s, int32_t* flat);
int     t1_replay_exhausted(const t1_sim* s);

/* contact flavour (monte_carlo.h:157-195, 274-316, 343-355, 443-516, 525-688) */
void    t1_contacts_init(t1_sim* s, int32_t n_seg, int64_t c1_pop, int64_t c2_pop);
void    t1_area(const t1_sim* s, double* area);
int64_t t1_num_contact_sites(const t1_sim* s, int which);
void    t1_contact_sites(const t1_sim* s, int which, int32_t* ids);
void    t1_contact_iteration(t1_sim* s, double dt, int64_t* pop /* [n_seg] */, int64_t* curr /* [n_seg-1] */);
int64_t t1_track_particle(t1_sim* s, double dt, uint64_t gid, int64_t max_steps, double* path /* [max_steps][3] */,
                          int32_t* reached); /* monte_carlo.h:786-818 */

#ifdef __cplusplus
}
#endif
#endif
