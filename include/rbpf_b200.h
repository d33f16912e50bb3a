/*
 * rbpf_b200.h -- C ABI of the B200-native RBPF per-scan update.
 *
 * Drop-in boundary for the hot path of amansanghvi/Thesis (motion -> scan match
 * -> likelihood weighting -> log-odds ray-cast -> systematic resampling).  The
 * reference has no FFI of its own for this path: it is driven through Python
 * duck-typed objects (Robot, robot.py:19-157; Map ABC, Map.py:16-102) from three
 * list comprehensions and one function in main.py:144,157-160, and its only
 * foreign call is the MATLAB engine RPC eng.matchScanCustom(...)
 * (hybridmap.py:244-251).  Each entry point below names the reference interface
 * it replaces; thesis_b200/particles.py is the ctypes host side that keeps the
 * reference's Python signatures on top.
 *
 * Conventions
 *   - every function returns 0 on success or a negative rbpf_status; the text of
 *     the last error of a handle is rbpf_last_error(h).
 *   - host pointers are caller-owned, read during the call (inputs) or written
 *     before return (outputs).  Pointers documented as DEVICE pointers must be
 *     device memory on the handle's GPU.
 *   - a handle is not thread-safe; all work is enqueued on the stream given at
 *     creation (0 = the legacy default stream) and calls that return data to the
 *     host synchronise that stream.
 *   - there is no CPU fallback: without a CUDA device rbpf_create fails with
 *     RBPF_ERR_CUDA.
 *   - poses are (x, y, theta) float64, covariances row-major 3x3 float64, maps
 *     are int8 log-odds in tenths (reference float64 value = tenths / 10, exact
 *     to 1.1e-14).
 */
#ifndef RBPF_B200_H
#define RBPF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rbpf_ctx *rbpf_handle;

typedef enum {
    RBPF_OK = 0,
    RBPF_ERR_ARG = -1,        /* bad argument / configuration */
    RBPF_ERR_CUDA = -2,       /* CUDA runtime failure (text in rbpf_last_error) */
    RBPF_ERR_POOL = -3,       /* tile pool exhausted */
    RBPF_ERR_RESAMPLE = -4,   /* main.py:66-67 AssertionError("Incorrect number of resampled weights.") */
    RBPF_ERR_WORLD = -5       /* a ray left the configured world extent (cells were dropped) */
} rbpf_status;

/* Motion-model families of the reference's IMU loaders (robot.py:45-57). */
typedef enum {
    RBPF_MOTION_ABSOLUTE = 0, /* IntelIMUData.py:22-36     u = (x, y, theta) */
    RBPF_MOTION_VELOCITY = 1, /* IntelRawIMUData.py:33-55 (Aces, Freid*, Obero, Bele) u = (vx, vy, w),
                                 par = (a_xy, b_xy, a_th, b_th) of get_cov_input_uncertainty */
    RBPF_MOTION_UNICYCLE = 2  /* DefaultIMUData.py:25-54   u = (v, w) */
} rbpf_motion_family;

/* rbpf_config.flags */
#define RBPF_FLAG_NDT_REFINE 1 /* run the NDT refinement stage of the reference matcher after the grid search
                                  (matchScanCustom.m:32-50: matchScans(..., 'MaxIterations', 500, 'CellSize', 0.1),
                                  accepted iff the refined pose passes isValidPose and 2*ndtScore > gridScore) */

typedef struct {
    int32_t n_particles;      /* particles owned by THIS handle (rank-local slice), main.py:44,87 */
    int32_t n_beams;          /* beams per sweep, <= 384 (180 Intel/ACES, 360 Freiburg, 361 UNSW/Bele) */
    int32_t n_samples;        /* proposal samples per particle, robot.py:17 NUM_SAMPLE_POINTS = 30; <= 32 */
    int32_t world_tiles_x;    /* odd number of 40 m reference tiles along x, centred on tile (0,0); x*y <= 64 */
    int32_t world_tiles_y;
    uint32_t pool_subtiles;   /* capacity of the copy-on-write pool in 160x160-cell sub-tiles (25,600 B each) */
    int32_t device;           /* CUDA device ordinal */
    int32_t rank;             /* rank / world of the particle sharding (0 / 1 on a single GPU) */
    int32_t world;
    int32_t flags;            /* RBPF_FLAG_* */
    uint64_t stream;          /* cudaStream_t to enqueue on, 0 = legacy default stream */
    uint64_t seed;            /* Philox key for device-side draws when the caller supplies none */
} rbpf_config;

typedef struct {
    uint32_t pool_subtiles;   /* capacity */
    uint32_t pool_in_use;     /* sub-tiles currently referenced */
    uint64_t cow_copies;      /* sub-tiles copied because they were shared (cumulative) */
    uint64_t fresh_allocs;    /* zero-filled sub-tiles handed out (cumulative) */
    uint64_t cells_dropped;   /* ray cells outside the world extent (cumulative) */
    uint64_t resamples;       /* scans whose resample triggered (cumulative) */
    uint64_t match_failed;    /* particle-scans whose match failed the isValidPose gate (cumulative) */
    uint64_t shared_refs;     /* page-table entries whose sub-tile is shared by >1 particle (snapshot) */
    uint64_t total_refs;      /* allocated page-table entries over all particles (snapshot) */
    uint64_t refcount_sum;    /* sum of sub-tile reference counts (snapshot; must equal total_refs) */
    uint64_t match_visits;    /* points visited by those passes (an aborted pass visits fewer than its points) */
    uint64_t match_points;    /* matcher points summed over all matches (exhaustive search = 231 * match_points visits) */
    uint64_t match_runs;      /* matcher searches actually run: duplicates of the last resample are bit-identical
                                 until the next weight stage and take their representative's result */
    uint64_t match_evals;     /* bitmap scoring passes of the matcher (group bounds + member rotations; exhaustive = 231 per match) */
    uint64_t ndt_evals;       /* NDT score evaluations of the refinement stage (cumulative, searches actually run) */
    uint64_t ndt_accepted;    /* searches whose refined pose replaced the grid pose (matchScanCustom.m:38-41) */
    uint64_t match_failed_zero; /* of match_failed: the optimum was the zero correction, which isValidPose rejects
                                   (matchScanCustom.m:55); the others sit on the border of the search window (:53-54) */
} rbpf_stats_t;

/* Replaces `particles = [Robot(eng) for _ in range(NUM_PARTICLES)]` (main.py:87,
 * Robot.__init__ robot.py:20-28, HybridMap.__init__ hybridmap.py:66-70): N
 * particles at pose 0, zero covariance, weight 1.0, one blank tile at (0,0). */
int rbpf_create(const rbpf_config *cfg, rbpf_handle *out);
int rbpf_destroy(rbpf_handle h);
const char *rbpf_last_error(rbpf_handle h);

/* Replaces `Lidar.__getitem__` -> `Scan.__init__` (lidar.py:28-31,76-80): one
 * sweep, ranges[B] in metres and angles[B] in radians, host pointers. */
int rbpf_set_scan(rbpf_handle h, const double *ranges, const double *angles, int32_t n_beams);

/* Replaces `[p.imu_update(reading) for p in particles]` (main.py:144,
 * robot.py:45-57).  u[4], par[4] (unused entries 0), dt in seconds. */
int rbpf_motion(rbpf_handle h, int32_t family, const double *u, double dt, const double *par);

/* Replaces the matcher half of Robot.map_update (robot.py:62-71):
 * HybridMap.get_scan_match (hybridmap.py:210-261) + eng.matchScanCustom
 * (matchScanCustom.m:1-58).  Results stay on the device (see getters). */
int rbpf_scan_match(rbpf_handle h);
/* Scan-to-previous-scan variant used on "adj" frames (main.py:156-159):
 * HybridMap.get_scan_adj (hybridmap.py:147-191).  last_scan_xy = n_points global
 * endpoints (x, y interleaved, host) of the previous scan, shared by all
 * particles (main.py:168). */
int rbpf_scan_match_adj(rbpf_handle h, const double *last_scan_xy, int32_t n_points);

/* Replaces the sampling/weighting half of Robot.map_update (robot.py:73-114):
 * proposal samples, _generate_sample_weight (robot.py:118-139), moments, weight
 * accumulation.  z = N*K*3 standard normals (host, particle-major) or NULL for
 * device Philox draws.  Particles whose match failed are left for
 * rbpf_integrate (their weight needs the updated map, robot.py:75-77). */
int rbpf_weight(rbpf_handle h, const double *z);
/* The same stage with the proposal samples themselves supplied by the caller: guesses = N*K*3 (host,
 * particle-major; rows of particles whose match failed are ignored).  This is the parity seam for
 * `guesses = np.random.multivariate_normal(scan_pose, scan_cov, 30)` (robot.py:81): the Python side draws
 * with NumPy itself, from the matcher results of rbpf_get_match, so the run consumes the global RNG stream
 * exactly like the reference and sees bit-identical samples. */
int rbpf_weight_guesses(rbpf_handle h, const double *guesses);

/* Replaces HybridMap.update (hybridmap.py:95-145) at every particle's current
 * pose, then the NaN-covariance weight fallback (robot.py:76-77) for particles
 * whose last match failed.  Also the map seeding of main.py:89-90 when called
 * before any match (fallback_weights = 0). */
int rbpf_integrate(rbpf_handle h, int32_t fallback_weights);

/* Replaces `particles = resample(particles)` (main.py:46-79,160) and the deep
 * copies of Robot.copy (robot.py:141-149) by copy-on-write page-table sharing.
 * u01 = the uniform of main.py:59 or NULL (device Philox).  ancestors_out
 * (nullable, host, N_global int32) and did_resample (nullable) synchronise. */
int rbpf_resample(rbpf_handle h, const double *u01, int32_t *ancestors_out, int32_t *did_resample);

/* One lidar event for throughput runs: set_scan + scan_match + weight +
 * integrate + resample with device-side draws and no host synchronisation. */
int rbpf_step(rbpf_handle h, const double *ranges, const double *angles, int32_t n_beams);

/* Per-stage device timing of rbpf_step with CUDA events on the handle's stream
 * (bench.py's roofline figures).  Enable for up to max_steps steps; read sums
 * ms[8] = set_scan, match, weight, raycast_prepare, raycast_cast,
 * weight_fallback, resample_plan, resample_apply and the number of steps. */
int rbpf_timing_enable(rbpf_handle h, int32_t max_steps);
int rbpf_timing_read(rbpf_handle h, double *ms_out8, int32_t *steps_out);

/* State access (host pointers; each call synchronises the stream). */
int rbpf_get_poses(rbpf_handle h, double *out_n3);        /* Robot.get_latest_pose robot.py:42-43 */
int rbpf_get_covs(rbpf_handle h, double *out_n9);         /* Robot._cov */
int rbpf_get_weights(rbpf_handle h, double *out_n);       /* Robot.weight()[-1] robot.py:39-40 */
int rbpf_set_poses(rbpf_handle h, const double *in_n3);
int rbpf_set_covs(rbpf_handle h, const double *in_n9);
int rbpf_set_weights(rbpf_handle h, const double *in_n);
/* Matcher results of the last rbpf_scan_match: pose[N*3], cov[N*9] (NaN when
 * invalid), score[N], valid[N], best[N*4] = (i, j, k, n_curr_points). */
int rbpf_get_match(rbpf_handle h, double *pose_n3, double *cov_n9, double *score_n, int32_t *valid_n, int32_t *best_n4);
/* NDT stage of the last match (matchScanCustom.m:32-50): out[N*2] = (score evaluations, 1 if the
 * refined pose replaced the grid pose).  Zeros when the stage is off or the grid match was invalid. */
int rbpf_get_match_refine(rbpf_handle h, int32_t *out_n2);
/* Switch the NDT stage on or off for the following matches (initially rbpf_config.flags & RBPF_FLAG_NDT_REFINE). */
int rbpf_set_refine(rbpf_handle h, int32_t on);
/* Running sum of the adjusted weights of the last triggered resample (main.py:57,62),
 * n_global doubles; for parity checks of the float64 summation order. */
int rbpf_get_resample_cumsum(rbpf_handle h, double *out_n_global);
/* Inject matcher results (pose[N*3], cov[N*9], valid[N]) in place of
 * rbpf_scan_match -- the seam at which the reference calls MATLAB
 * (hybridmap.py:244-256); lets the weighting stage be checked against the
 * reference with a canned matcher answer. */
int rbpf_set_match(rbpf_handle h, const double *pose_n3, const double *cov_n9, const int32_t *valid_n);
/* Score slice (29x29 int32, row j, column i) at the best rotation of one particle: re-runs the
 * last match (scan-to-map or scan-to-previous-scan) for that particle. */
int rbpf_get_match_slice(rbpf_handle h, int32_t particle, int32_t *out_29x29);

/* One 40 m reference tile of one particle as the reference stores it: 800x800
 * float64 [ix][iy] (gridmap.py:32).  *exists = 0 when the particle has no such
 * HybridMapEntry (out is zero-filled).  cx, cy = tile centre in metres. */
int rbpf_export_tile(rbpf_handle h, int32_t particle, int32_t cx, int32_t cy, double *out_800x800, int32_t *exists);
/* Centres of the particle's existing tiles: out_xy[2*max_tiles]; returns count in *n. */
int rbpf_list_tiles(rbpf_handle h, int32_t particle, int32_t *out_xy, int32_t max_tiles, int32_t *n);

/* HybridMap.get_occupied_points (hybridmap.py:303-313; main.py:171,176 plots it):
 * cells with log-odds > 1.0 of one particle, thresholded and compacted on the
 * device, in the reference's cell units.  out_xy may be NULL (count only). */
int rbpf_occupied_points(rbpf_handle h, int32_t particle, double *out_xy, int64_t max_points, int64_t *n);
/* Checkpoint / resume of the whole particle set (the reference shelves particle 0
 * only, main.py:183-210): state, page tables and every sub-tile in use.  The
 * reading handle must have the same configuration. */
int rbpf_checkpoint_write(rbpf_handle h, const char *path);
int rbpf_checkpoint_read(rbpf_handle h, const char *path);

/* The same on the device: rbpf_snapshot copies the complete mutable state (pool, reference counts, page
 * tables, poses, covariances, weights) into shadow buffers -- allocated on first use, doubling the handle's
 * memory -- and rbpf_restore copies it back, both stream-ordered.  Lets a caller rewind the filter, e.g. to
 * run the same scans twice (bench.py: device-resident loop and host-buffer loop on identical work). */
int rbpf_snapshot(rbpf_handle h);
int rbpf_restore(rbpf_handle h);

/* Device-side errors are sticky and deferred: the kernels set a flag (pool exhausted -> RBPF_ERR_POOL, the
 * reference's resample assertion main.py:66-67 -> RBPF_ERR_RESAMPLE, internal bound -> RBPF_ERR_WORLD) and go
 * on without corrupting state -- after pool exhaustion scans are no longer integrated and resampling is
 * frozen.  The status is returned by the next synchronising call (rbpf_synchronize, getters, rbpf_resample with
 * outputs) and by rbpf_step at most two steps later, and keeps being returned until rbpf_clear_errors. */
int rbpf_clear_errors(rbpf_handle h);
int rbpf_stats(rbpf_handle h, rbpf_stats_t *out);
/* Where the matcher kernel (the MATLAB matchScanCustom call of hybridmap.py:244-251) spends its time:
 * SM clocks since creation, out[16].  Slots 0-8 are summed over CTAs (one search each, thread 0 between
 * barriers): 0 frame + curr points, 1 map gather + threshold, 2 both dilations, 3 reference-set mask
 * (hybridmap.py:230-239), 4 seed rotations + group bounds, 5 group ranking, 6 member rotations, 7 covariance,
 * 8 NDT stage.
 * Slots 10-12 are summed over warps (busy time, without barrier waits): seeds, group bounds, member
 * rotations; 13-15 are the points those three phases visited. */
int rbpf_match_phase_clocks(rbpf_handle h, uint64_t *out16);
int rbpf_synchronize(rbpf_handle h);

/* Constants of the restated matcher, for callers that need the lattice. */
double rbpf_rot_step(void);
int32_t rbpf_rot_count(void);

/* ---- multi-GPU resampling (one handle per rank; the caller moves the bytes
 * with NCCL, see thesis_b200/dist.py).  Pointers here are DEVICE pointers. */

/* Local weights of this rank, contiguous N doubles, to be all-gathered. */
int rbpf_weights_device_ptr(rbpf_handle h, uint64_t *dev_ptr);
/* Global systematic resample (main.py:46-67) on the all-gathered weights of all
 * ranks (n_global doubles, identical on every rank): fills the handle's ancestor
 * vector, moves nothing.  ancestors_out (nullable host, n_global) synchronises. */
int rbpf_resample_global(rbpf_handle h, uint64_t weights_all_dev, int32_t n_global, const double *u01,
                         int32_t *did_resample, int32_t *ancestors_out);
/* Sender: src_slots[n] = local particles a peer needs (each once).  _count
 * de-duplicates their sub-tiles and reports how many travel and the packed size;
 * _pack writes [records | page tables | sub-tile payloads] to dev_buf. */
int rbpf_migrate_count(rbpf_handle h, const int32_t *src_slots, int32_t n, int32_t *n_subtiles, int64_t *bytes);
int rbpf_migrate_pack(rbpf_handle h, uint64_t dev_buf);
int64_t rbpf_migrate_bytes(rbpf_handle h, int32_t n_particles, int32_t n_subtiles);
/* Local half of the resample: gather local ancestors, fix reference counts
 * (replaces Robot.copy robot.py:141-149 by page-table sharing).  Call after the
 * packs and before the unpacks. */
int rbpf_resample_apply_local(rbpf_handle h);
/* The same with the reference-count pass deferred: the gather of the local ancestors
 * (main.py:69-76 by page-table sharing) runs now; the reference counts -- and with them
 * every free, and every in-place write into a tile a peer may still be pulling -- are
 * launched by the next call that needs the pool's bookkeeping (rbpf_integrate, a
 * resample, statistics, snapshots, rbpf_synchronize ...), behind `gate_event`
 * (a cudaEvent_t recorded by the caller once every peer has finished pulling: the
 * job-wide barrier, taken off the critical path; 0 = no gate).  Motion, matching and
 * weighting of the next scan run in between: they only read tiles. */
int rbpf_resample_apply_local_deferred(rbpf_handle h, uint64_t gate_event);

/* Pull migration over peer memory (NVLink): instead of pack -> NCCL -> unpack, the
 * receiving rank maps the source rank's particle state with CUDA IPC once and then
 * reads page tables, poses and sub-tiles straight into their final place.
 * rbpf_peer_export   describes this handle's buffers (to be all-gathered as bytes);
 * rbpf_peer_attach   maps a peer's buffers (same process: the raw pointers are used);
 * rbpf_migrate_pull  every local slot whose ancestor (last rbpf_resample_global) lives on
 *                    another rank becomes a copy of that particle.  The plan is the
 *                    ancestor vector on the device: stream-ordered, no host copy, no
 *                    host synchronisation.  The caller orders the steps of one resample
 *                    as: all ranks pull -> job-wide barrier -> rbpf_resample_apply_local
 *                    -> rbpf_resample_commit, because a source may only free or overwrite
 *                    state after every peer has read it (or: pull -> barrier on a side
 *                    stream + event -> rbpf_resample_apply_local_deferred(event) -> commit). */
typedef struct {
    unsigned char ipc[9][64]; /* cudaIpcMemHandle_t of pool, page tables x2, poses x2, covariances x2, tile masks x2 */
    uint64_t ptr[9];          /* the same allocations as device pointers of the exporting process */
    int64_t pid;              /* exporting process */
    int32_t parity;           /* which of the double buffers is current */
    int32_t device;
    int32_t n_particles;
    uint32_t pool_subtiles;
    int32_t nsub;
    int32_t reserved;
} rbpf_peer_view;
int rbpf_peer_export(rbpf_handle h, rbpf_peer_view *out);
int rbpf_peer_attach(rbpf_handle h, int32_t peer_rank, const rbpf_peer_view *view);
int rbpf_migrate_pull(rbpf_handle h);
/* The same with the sub-tile payloads copied on `copy_stream` (a cudaStream_t; 0 = the handle's stream, i.e.
 * rbpf_migrate_pull): claims, allocation, page tables and particle state stay on the handle's stream, so the rest of
 * the resample and the next scan go ahead; rbpf_scan_match matches the local particles first and the migrated ones
 * behind the copies, every other call that touches tiles waits for them.  The caller's job-wide barrier has to
 * follow the copies, i.e. be issued on `copy_stream`. */
int rbpf_migrate_pull_async(rbpf_handle h, uint64_t copy_stream);
/* Receiver: adopt a packed buffer; local slot dst_slots[i] becomes a copy of
 * received record rec_index[i] (weight 1.0, main.py:77-78). */
int rbpf_migrate_unpack(rbpf_handle h, uint64_t dev_buf, int32_t n_particles, int32_t n_subtiles,
                        const int32_t *dst_slots, const int32_t *rec_index, int32_t m);
/* Make the resampled particle buffers current. */
int rbpf_resample_commit(rbpf_handle h);

#ifdef __cplusplus
}
#endif
#endif /* RBPF_B200_H */
