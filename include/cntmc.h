/* cntmc.h -- C ABI of libcntmc.so, the B200-native exciton hop engine.
 *
 * The reference (amirhosseindavoody/cnt_film_monte_carlo) has no plugin or FFI boundary: main() calls C++ methods
 * on one mc::monte_carlo object (src/main.cpp:64-76, 85-105).  This header is the boundary a maintainer binds
 * instead: one opaque handle per simulation, one entry point per mc::monte_carlo method on the hot path.  Each entry
 * point names the reference interface it replaces (file:line into /root/reference/src).  INTEGRATION.md shows the
 * C++ shim (class mc::monte_carlo re-implemented over these calls) and the ctypes binding.
 *
 * Conventions: every function returning int gives 0 on success and a negative code on failure; the message is
 * available from cntmc_last_error(h) (or cntmc_last_error(NULL) when no handle exists yet).  No exception crosses the
 * boundary.  The caller owns every host buffer it passes; the handle owns all device memory.  One host thread per
 * handle; a cntmc_t drives one GPU, a cntmc_multi_t (end of this header) 1..8 GPUs of one box.  All pointers are
 * plain host pointers unless the name says "dev".
 * There is no CPU execution path: without a CUDA device cntmc_kubo_init() fails with CNTMC_ERR_CUDA.
 */
#ifndef CNTMC_H
#define CNTMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cntmc_handle cntmc_t;

enum {
  CNTMC_OK = 0,
  CNTMC_ERR_INVALID = -1, /* bad argument / bad JSON / call out of order (std::invalid_argument in the reference) */
  CNTMC_ERR_CUDA = -2,    /* a CUDA runtime call failed, or no device */
  CNTMC_ERR_STATE = -3,   /* simulation reached a state the reference leaves undefined (e.g. a site with no neighbour,
                             scatterer.h:91) or a bounded device loop hit its guard */
  CNTMC_ERR_REPLAY = -4   /* a replayed draw list ran out */
};

/* ---- life cycle ------------------------------------------------------------------------------------------------- */

/* monte_carlo::monte_carlo(const nlohmann::json&)  monte_carlo/monte_carlo.h:116-136  (+ main.cpp:41-54).
 * json_text is either a whole input.json or its "exciton monte carlo" block.  Directory preparation (output
 * rotation, prepare_directory.hpp) is left to the host shim; the handle itself touches no files except the mesh. */
int  cntmc_create(const char* json_text, cntmc_t** out);
void cntmc_destroy(cntmc_t* h);
const char* cntmc_last_error(const cntmc_t* h);
const char* cntmc_version(void);

/* CUDA placement (call before *_init).  device < 0 keeps the current device; stream is a cudaStream_t (NULL = the
 * legacy default stream).  All kernels and copies of the handle are issued on that stream. */
int cntmc_set_device(cntmc_t* h, int device);
int cntmc_set_stream(cntmc_t* h, void* cuda_stream);

/* ---- inputs ----------------------------------------------------------------------------------------------------- */

/* monte_carlo::create_scatterers  monte_carlo.h:199-271: reads the six single_cnt.{pos,orient}.{x,y,z}.dat files.
 * dir == NULL uses "mesh input directory" of the JSON ('~' expanded like prepare_directory.hpp:12-16). */
int cntmc_load_mesh(cntmc_t* h, const char* dir);
/* same, from memory: pos_nm and orient are [3][n_tubes*n_cols] (x plane, y plane, z plane; tube-major), positions in nm */
int cntmc_set_mesh(cntmc_t* h, int64_t n_tubes, int64_t n_cols, const double* pos_nm, const double* orient);

/* scattering_struct  monte_carlo/scattering_struct.h:11-35: install a precomputed 4-D table (e.g. a "davoody" table
 * saved by scattering_struct::save, :56-94) instead of building the closed-form one.  rates is [theta][z][a1][a2]. */
int cntmc_set_rate_table(cntmc_t* h, const int32_t dims[4], const double* theta, const double* z_shift,
                         const double* axis_shift_1, const double* axis_shift_2, const double* rates);
int cntmc_get_rate_table_dims(const cntmc_t* h, int32_t dims[4]);
int cntmc_get_rate_table(const cntmc_t* h, double* theta, double* z_shift, double* axis_shift_1, double* axis_shift_2,
                         double* rates);

/* scattering_struct::save  scattering_struct.h:56-94: dir/scat_table.{theta,z_shift,axis_shift_1,axis_shift_2,rates}.dat
 * (same files and layout; 17 significant digits so that a saved table loads back bit for bit), and the loader the
 * reference lacks (it only writes the files; visualization/monte_carlo_results.py:261-282 reads them the same way). */
int cntmc_save_rate_table(const cntmc_t* h, const char* dir);
int cntmc_load_rate_table(cntmc_t* h, const char* dir);

/* ---- Green-Kubo flavour ------------------------------------------------------------------------------------------- */

/* monte_carlo::kubo_init  monte_carlo/monte_carlo.cpp:254-305: rate table (create_scattering_table :24-61 for
 * forster/wong), trim_scats, find_simulation_domain, create_scatterer_buckets, set_max_rate (here: the CSR
 * neighbour-table build on the GPU), injection_region, get_removal_domain. */
int cntmc_kubo_init(cntmc_t* h);

/* monte_carlo::kubo_create_particles  monte_carlo.cpp:308-316.  n_particles <= 0 uses "number of particles for kubo
 * simulation".  Exciton i gets the counter-based stream (seed, first_global_id + i): results do not depend on how a
 * population is split over GPUs. */
int cntmc_kubo_create_particles(cntmc_t* h, int64_t n_particles, uint64_t seed, uint64_t first_global_id);
/* same, but every draw comes from a recorded list (the reference's own rand() stream, split per exciton):
 * exciton i consumes draws[offsets[i] .. offsets[i+1]).  logs (may be NULL) holds log(draw/RAND_MAX) as computed by
 * the host libm for each draw, which makes free-flight times bit-identical to a glibc run. */
int cntmc_kubo_create_particles_replay(cntmc_t* h, int64_t n_particles, const int64_t* offsets, const int32_t* draws,
                                       const double* logs);

/* nsteps x { monte_carlo::kubo_step(dt)  monte_carlo.cpp:319-342 ; the ensemble averages written by
 * kubo_save_avg_dispalcement_squared  :396-406 }.  msd_out (may be NULL) receives [nsteps][3] = <dx^2>,<dy^2>,<dz^2>
 * after each step, averaged over this handle's excitons. */
int cntmc_kubo_step(cntmc_t* h, double dt, int64_t nsteps, double* msd_out);
/* same, but leaves the un-normalised sums on the device for a collective: dev_sums is a device pointer to
 * [nsteps][4] doubles = sum dx^2, sum dy^2, sum dz^2, hops in that step.  Asynchronous on the handle's stream. */
int cntmc_kubo_step_dev(cntmc_t* h, double dt, int64_t nsteps, double* dev_sums);
/* same as cntmc_kubo_step for a population that lives in HOST memory (like the reference's std::vector<particle>):
 * uploads the state, steps, downloads it again.  Arrays as in cntmc_get_particles. */
int cntmc_kubo_step_host_state(cntmc_t* h, double dt, int64_t nsteps, int64_t n_particles, int32_t* site, double* pos,
                               double* delta, double* ff, uint8_t* heading, uint32_t* ndraw, double* msd_out);

double  cntmc_time(const cntmc_t* h);              /* monte_carlo::time            monte_carlo.h:139 */
double  cntmc_kubo_max_time(const cntmc_t* h);     /* monte_carlo::kubo_max_time   monte_carlo.h:833 */
double  cntmc_time_step(const cntmc_t* h);         /* "monte carlo time step"      main.cpp:62 */
int64_t cntmc_number_of_particles(const cntmc_t* h); /* monte_carlo::number_of_particles monte_carlo.h:154 */
int64_t cntmc_hops(const cntmc_t* h);              /* scattering events so far (the metric's unit) */
int64_t cntmc_reinjections(const cntmc_t* h);
/* bookkeeping behind the roofline's algorithmic bytes: chain sites crossed in flight and cumulative-rate entries probed */
int64_t cntmc_crossings(const cntmc_t* h);
int64_t cntmc_probes(const cntmc_t* h);

/* ---- contact flavour ------------------------------------------------------------------------------------------------ */

/* monte_carlo::init  monte_carlo.h:157-195 (contacts = first and last of "number of segments" slabs along y;
 * create_particles :274-316 with a linear profile from c1_pop to c2_pop; the reference hard-codes 1100 and 0). */
int cntmc_init(cntmc_t* h, int64_t c1_pop, int64_t c2_pop, uint64_t seed, int64_t capacity);
/* same, with every draw taken from recorded lists (the reference's own rand() stream, split per exciton): exciton id g --
 * ids in order of birth: the initial population in creation order, then per iteration the excitons repopulate creates on
 * contact 1, then on contact 2 (monte_carlo.h:443-491) -- consumes draws[offsets[g] .. offsets[g+1]).  logs as in
 * cntmc_kubo_create_particles_replay.  n_ids must cover every exciton created while stepping. */
int cntmc_init_replay(cntmc_t* h, int64_t c1_pop, int64_t c2_pop, int64_t n_ids, const int64_t* offsets, const int32_t* draws,
                      const double* logs);
/* ids of the excitons alive, in the order of cntmc_get_particles (contact mode) */
int cntmc_get_gids(const cntmc_t* h, uint64_t* gid);
/* nsteps x { monte_carlo::step(dt) :343-355 ; save_population_profile :566-573 ; save_currents :593-636 ;
 * repopulate_contacts :443-455 }.  pop_out [nsteps][n_seg] excitons per slab, curr_out [nsteps][n_seg-1] net
 * crossings per interface (raw counts; the shim divides by area*dy and area*dt like the reference's writers). */
int cntmc_step(cntmc_t* h, double dt, int64_t nsteps, int64_t* pop_out, int64_t* curr_out);
int cntmc_step_dev(cntmc_t* h, double dt, int64_t nsteps, int64_t* dev_bins /* [nsteps][2*n_seg-1] */);
int cntmc_get_area(const cntmc_t* h, double* area /* [n_seg] */);                /* monte_carlo::get_area :646-688 */
int cntmc_num_contact_sites(const cntmc_t* h, int which, int64_t* n);           /* monte_carlo::contact_scats :494-516 */
int cntmc_get_contact_sites(const cntmc_t* h, int which, int32_t* ids);
int cntmc_number_of_segments(const cntmc_t* h);
/* monte_carlo::get_scatterer_statistics  monte_carlo.h:691-719: sites per slab, pop [n_seg] (the shim writes
 * scatterer_statistics.dat from it) */
int cntmc_get_scatterer_statistics(const cntmc_t* h, int64_t* pop);
/* monte_carlo::track_particle  monte_carlo.h:786-818: one exciton born on the first contact is stepped until it enters
 * the last slab (or max_steps, which the reference does not have); path [max_steps][3] receives its position after
 * every step, *n_steps the rows written, *reached (may be NULL) whether the last slab was entered.  The exciton's
 * draws come from the stream (seed, global_id), or from a recorded list when n_replay > 0 (replay_logs may be NULL). */
int cntmc_track_particle(cntmc_t* h, double dt, uint64_t seed, uint64_t global_id, int64_t n_replay,
                         const int32_t* replay_draws, const double* replay_logs, int64_t max_steps, double* path,
                         int64_t* n_steps, int32_t* reached);

/* ---- read-back (parity tests, checkpoints, output writers) ------------------------------------------------------------ */

int cntmc_num_sites(const cntmc_t* h, int64_t* n);
/* post-trim site list: pos/orient [3][N] in metres, chain links, Gamma_i (scatterer::_max_rate) and its inverse */
int cntmc_get_sites(const cntmc_t* h, double* pos, double* orient, int32_t* left, int32_t* right, double* max_rate,
                    double* inv_max_rate);
int cntmc_get_domain(const cntmc_t* h, double lo_hi[6]);          /* find_simulation_domain monte_carlo.h:328-340 */
int cntmc_get_removal_domain(const cntmc_t* h, double lo_hi[6]);  /* get_removal_domain monte_carlo.cpp:231-251 */
int cntmc_num_inject(const cntmc_t* h, int64_t* n);               /* injection_region monte_carlo.cpp:203-228 */
int cntmc_get_inject(const cntmc_t* h, int32_t* ids);
/* the neighbour table = scatterer::find_neighbors (scatterer.cpp:34-83) of every site, as CSR */
int cntmc_csr_nnz(const cntmc_t* h, int64_t* nnz);
int cntmc_get_csr(const cntmc_t* h, int64_t* row_ptr /* [N+1] */, int32_t* nbr /* [nnz] */, double* cum /* [nnz] */);
/* one row of the table (tables of 1e9 entries are not read back whole): up to cap entries, *len = the row's length */
int cntmc_get_csr_row(const cntmc_t* h, int64_t site, int64_t cap, int32_t* nbr, double* cum, int64_t* len);
/* rows holding a pair whose theta fell within 1e-9 grid pitches of a grid midpoint, where the device's acos could pick
 * another table index than glibc's: every such row was recomputed on the host with glibc's acos and patched in
 * (expected count: 0; option guard_ppb widens the band, for tests) */
int64_t cntmc_csr_midpoint_guards(const cntmc_t* h);
double  cntmc_csr_build_seconds(const cntmc_t* h);

/* exciton state: site [P], pos [3][P], delta [3][P] (particle::_delta_pos), ff [P], heading [P], ndraw [P]; any
 * pointer may be NULL */
int cntmc_get_particles(const cntmc_t* h, int32_t* site, double* pos, double* delta, double* ff, uint8_t* heading,
                        uint32_t* ndraw);

/* record the site reached by each scattering event of the next cntmc_kubo_step call (tests only; cap events per
 * exciton).  After the step: counts [P] and sites [P][cap]. */
int cntmc_trace_enable(cntmc_t* h, int32_t cap);
int cntmc_trace_get(const cntmc_t* h, int32_t* counts, int32_t* sites);

/* ---- tuning (never changes results) ------------------------------------------------------------------------------------
 * chunk_steps  time steps per hop-kernel launch (default 64)          stage_mb     cap on the staging buffer in MiB (0 = auto)
 * hot_pct      share of the lane blocks that serve the most active excitons first (default 25; re-tuned for 7 blocks per SM: 22-26 within 1 %, 20 and 35 lose 3-8 %)
 * deep_thr     Gamma*dt from which an exciton is handed to the trap solver, a second kernel that walks trapped excitons
 *              from a register-resident window of site records (default 0 = off: bit-identical results, but slower on
 *              every workload measured so far); deep_blocks (blocks per SM of its launch, default 4), deep_rounds
 *              (2: it hands excitons that left their trap back to the lanes once per launch; default 2); deep_group (8: the
 *              window walk; 1: the ordinary loop, one exciton per lane, in warps that hold trapped excitons only) with
 *              trap_burst (deep_group 1: events a lane may run in a row, default 1), deep_overlap (deep_group 1: the trap
 *              kernel runs beside the lane kernel of the same launch on a stream of its own and takes the deferred excitons
 *              while they arrive) and overlap_trap_blocks (its blocks per SM, 1..4).  All measured, none a gain:
 *              profiles/round2_trap_solver.txt
 * csr_warp     1 (default): the table's fill pass runs one warp per row; 0: one thread per row (the cross-check)
 * host_slices  cntmc_kubo_step_host_state steps the uploaded population in this many slices on their own streams so that
 *              the copies of one overlap the kernels of the others (default 4; populations below 65536 per slice: 1)
 * gid_base_shift56  contact mode: stream ids start at value * 2^56 (cntmc_multi keeps the GPUs' streams apart with it)
 * occupancy    resident 128-thread blocks per SM the hop kernel is compiled for (4 to 8; default 0 = 8 when the tables exceed the L2 cache or a Green-Kubo population has 2e6 excitons or more, else 7)
 * top_entries  1: the three widest entries of a row are tried before the row is searched (default)
 * runs         1: chain walks over memory-consecutive sites read segment times instead of chasing records (default)
 * dirs         1: last legs that leave from a site use stored unit vectors (default)
 * stats        1: run the instrumented hop kernel (probe / crossing counters for cntmc_probes / cntmc_crossings)
 * time_kernels 1: CUDA events around every hop-kernel launch (cntmc_last_kernel_ms) */
int cntmc_set_option(cntmc_t* h, const char* name, int64_t value);
int64_t cntmc_get_option(const cntmc_t* h, const char* name);
/* device time of the last kubo_step / step call's kernels in milliseconds (CUDA events on the handle's stream) */
double cntmc_last_step_ms(const cntmc_t* h);
int64_t cntmc_last_step_launches(const cntmc_t* h);
/* wait for the handle's stream, report errors of asynchronous (*_dev) calls, refresh the cumulative counters */
int cntmc_sync(cntmc_t* h);
/* with option "time_kernels" = 1: summed device time and count of the hop-kernel launches of the last step call */
double  cntmc_last_kernel_ms(const cntmc_t* h);
int64_t cntmc_last_kernel_launches(const cntmc_t* h);
/* self-test, no reference counterpart: the hop kernels take particle::fly's |pos - next.pos| (particle.cpp:39) and
 * normalise(next.pos - pos) (particle.cpp:47) with a square root and a division written out without the compiler's slow-path
 * calls (csrc/hop_core.h sqrt_walk / div3_walk).  Runs both on about n pseudo-random operands on `device` against the IEEE
 * operations: counts = {square roots that differ, quotients that differ, in-range operands flagged, out-of-range operands not
 * flagged}; all four must be zero.  Errors are reported by cntmc_last_error(NULL). */
int cntmc_dbg_walk_arith(int device, int64_t n, uint64_t seed, int64_t counts[4]);

/* ---- one simulation on several GPUs of one box ---------------------------------------------------------------------------
 * The reference parallelises monte_carlo::kubo_step / step over its particle list with OpenMP (monte_carlo.cpp:320-338,
 * monte_carlo.h:345-351).  A cntmc_multi_t does the same over GPUs: the read-only tables are replicated (each GPU
 * builds its own neighbour table), the excitons are split by global id (the streams are keyed by global id, so every
 * trajectory is the same bits for any number of GPUs), and each step call ends with ONE ncclAllReduce (sum) of the
 * per-step rows -- the sums of kubo_save_avg_dispalcement_squared (monte_carlo.cpp:396-400) or the population / current
 * bins (monte_carlo.h:566-573, 626-636) -- on every GPU's own stream.  NCCL is bound at run time (libnccl.so.2).
 * One host thread; devices == NULL means GPUs 0..n_devices-1. */
typedef struct cntmc_multi cntmc_multi_t;
int  cntmc_multi_create(const char* json_text, int n_devices, const int* devices, cntmc_multi_t** out);
void cntmc_multi_destroy(cntmc_multi_t* m);
const char* cntmc_multi_last_error(const cntmc_multi_t* m);
int  cntmc_multi_num_devices(const cntmc_multi_t* m);
cntmc_t* cntmc_multi_handle(cntmc_multi_t* m, int i);     /* GPU i's own handle: read-backs, per-GPU options */
int  cntmc_multi_nccl_version(void);                      /* ncclGetVersion of the library that was bound, -1 if none */
int  cntmc_multi_load_mesh(cntmc_multi_t* m, const char* dir);                                   /* cntmc_load_mesh on every GPU */
int  cntmc_multi_set_mesh(cntmc_multi_t* m, int64_t n_tubes, int64_t n_cols, const double* pos_nm, const double* orient);
int  cntmc_multi_set_option(cntmc_multi_t* m, const char* name, int64_t value);
int  cntmc_multi_kubo_init(cntmc_multi_t* m);                                                    /* monte_carlo::kubo_init */
/* monte_carlo::kubo_create_particles: n_particles excitons in total, shard r = ids [first_r, first_r + n_r) */
int  cntmc_multi_kubo_create_particles(cntmc_multi_t* m, int64_t n_particles, uint64_t seed);
/* nsteps x { kubo_step ; kubo_save_avg_dispalcement_squared }: msd_out [nsteps][3] averaged over the WHOLE population */
int  cntmc_multi_kubo_step(cntmc_multi_t* m, double dt, int64_t nsteps, double* msd_out);
/* exciton state of the whole population in global-id order (arrays as cntmc_get_particles, P = all excitons) */
int  cntmc_multi_get_particles(const cntmc_multi_t* m, int32_t* site, double* pos, double* delta, double* ff, uint8_t* heading,
                               uint32_t* ndraw);
int64_t cntmc_multi_number_of_particles(const cntmc_multi_t* m);
int64_t cntmc_multi_hops(const cntmc_multi_t* m);
double  cntmc_multi_time(const cntmc_multi_t* m);
/* monte_carlo::init / step+save_metrics+repopulate_contacts with the contact populations split over the GPUs; bins
 * summed over GPUs as in cntmc_step */
int  cntmc_multi_init(cntmc_multi_t* m, int64_t c1_pop, int64_t c2_pop, uint64_t seed);
int  cntmc_multi_step(cntmc_multi_t* m, double dt, int64_t nsteps, int64_t* pop_out, int64_t* curr_out);

/* ---- davoody rate table ("rate type":"davoody") -------------------------------------------------------------------------
 * The reference builds this table in monte_carlo::create_scattering_table (monte_carlo.cpp:24-49: one cnt object per
 * entry of the JSON's "cnts", cnt::calculate_exciton_dispersion each) and monte_carlo::create_davoody_scatt_table
 * (monte_carlo.cpp:64-153: exciton_transfer::first_order for every (theta, z shift, axis shift 1, axis shift 2) entry,
 * 21 x 11 x 11 x 11 = 27 951 placements in the shipped input.json).  Here: the tube physics runs on the host once per
 * tube, the placement-independent factors once per tube pair, and every table entry is one thread block of one kernel.
 * Errors of this section are reported by cntmc_davoody_last_error() (they have no cntmc_t). */
typedef struct cntmc_tube     cntmc_tube_t;
typedef struct cntmc_transfer cntmc_transfer_t;
const char* cntmc_davoody_last_error(void);
/* arma::eig_sym as cnt.cpp:950-962 uses it (eigenvalues ascending, eigenvectors in the columns of V; the reference's is LAPACK
 * zheevd behind Armadillo, this is csrc/herm_eig.h, cyclic Jacobi): a is n x n row-major {re, im} pairs, w [n], v like a.  Host
 * only; exported so that the CPU suite can pin the solver to LAPACK (tests/test_davoody_host.py) */
int cntmc_hermitian_eig(int n, const double* a_re_im, double* w, double* v_re_im);
/* measurement aid, no reference counterpart: FP64 fused multiply-add peak of a device in TFLOP/s (a register-only kernel), the
 * denominator of the table kernel's roofline in tools/davoody_bench.py */
int cntmc_fp64_peak(int device, double* tflops);

/* cnt::cnt(json, dir) + cnt::calculate_exciton_dispersion  exciton_transfer/cnt.h:159-193, cnt.cpp:1056-1081: chirality
 * (n, m), length in cnt unit cells ("length": [L, "cnt unit cells"]).  Host only; writes no files.  NULL on failure. */
cntmc_tube_t* cntmc_tube_create(int n, int m, int length_cells);
void          cntmc_tube_destroy(cntmc_tube_t* t);
/* ints = {n, m, cells, Nu, M, Q, nk (K2-extended), sites = Nu*cells}; reals = {cnt::radius() cnt.h:261, length_in_meter()
 * :318, Au() :346, seconds the build took} */
int cntmc_tube_info(const cntmc_tube_t* t, int32_t ints[8], double reals[4]);
/* cnt::A1() / A2_singlet() / A2_triplet()  cnt.h:273-309: which = 0 / 1 / 2.  dims = {nk_cm, n_principal, nk_c,
 * ik_cm_range[0]}; energy is exciton_struct::energy(ik_cm_idx, n), row-major [nk_cm][n_principal], joules */
int cntmc_tube_exciton_dims(const cntmc_tube_t* t, int which, int32_t dims[4]);
int cntmc_tube_exciton_energy(const cntmc_tube_t* t, int which, double* energy);

/* exciton_transfer::exciton_transfer(cnt1, cnt2)  exciton_transfer.h:37-48 (which fixes 300 K and 4 meV; here they are
 * arguments, broadening in joules): relevant states, matched pairs, Q of every pair, site phases; uploads them to
 * `device` (< 0: the current one).  Both tubes must outlive the transfer.  NULL on failure (no CUDA device included). */
cntmc_transfer_t* cntmc_transfer_create(const cntmc_tube_t* donor, const cntmc_tube_t* acceptor, double temperature_kelvin,
                                        double broadening_joule, int device);
void              cntmc_transfer_destroy(cntmc_transfer_t* x);
/* ints = {donor states, acceptor states, matched pairs, distinct donor K_cm, distinct acceptor K_cm, K_cm per pass,
 * threads per block, shared memory bytes}; reals = {temperature, broadening, last kernel ms, launches so far} */
int cntmc_transfer_info(const cntmc_transfer_t* x, int32_t ints[8], double reals[4]);
/* per matched pair, in the reference's pair order: Q (calculate_Q, exciton_transfer.cpp:274-304) as re,im; the thermal
 * prefactor (2 pi / hbar) exp(-E_i/kT)/Z and the lorentzian of first_order's sum (:431).  Any pointer may be NULL. */
int cntmc_transfer_pair_factors(const cntmc_transfer_t* x, double* q_re_im, double* boltzmann, double* lorentzian);
/* exciton_transfer::first_order(z_shift, {axis_shift_1, axis_shift_2}, theta)  exciton_transfer.cpp:395-441 for n
 * placements at once (theta in radians, lengths in metres); rate[n] in 1/s */
int cntmc_transfer_first_order(cntmc_transfer_t* x, int64_t n, const double* z_shift, const double* axis_shift_1,
                               const double* axis_shift_2, const double* theta, double* rate);
/* the loop nest of monte_carlo::create_davoody_scatt_table  monte_carlo.cpp:114-137 over the four axes; rates is
 * [theta][z][a1][a2] like cntmc_set_rate_table's */
int cntmc_transfer_table(cntmc_transfer_t* x, const int32_t dims[4], const double* theta, const double* z_shift,
                         const double* axis_shift_1, const double* axis_shift_2, double* rates);
/* the same, installed into a simulation as its scattering table (create_davoody_scatt_table's return value,
 * monte_carlo.cpp:139); must precede cntmc_kubo_init / cntmc_init like cntmc_set_rate_table */
int cntmc_create_davoody_table(cntmc_t* h, cntmc_transfer_t* x, const int32_t dims[4], const double* theta,
                               const double* z_shift, const double* axis_shift_1, const double* axis_shift_2);

#ifdef __cplusplus
}
#endif
#endif /* CNTMC_H */
