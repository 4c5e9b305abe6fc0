/* reinfocus_b200 C-ABI: the drop-in boundary of the B200-native reinfocus hot path.
 *
 * The reference (jeffwhunter/reinfocus) is pure Python + numba-CUDA and has no FFI of its
 * own; the calls below are what a ctypes binding inside the reference's
 * reinfocus/graphics/render.py and reinfocus/vision.py would bind instead of launching
 * numba kernels / calling OpenCV (INTEGRATION.md shows that binding). Each entry point
 * cites the reference interface it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types.
 *   - every function returns 0 on success or a negative rf_status; rf_last_error() gives
 *     the message of the last failure on that context (rf_last_global_error() for
 *     failures without a context, e.g. in rf_create).
 *   - "d_" pointers are device pointers on the context's GPU, owned by the caller (e.g.
 *     torch CUDA tensors); "h_" pointers are host pointers (pinned memory makes the
 *     copies asynchronous). The context owns only the RNG-state buffer, the packed scene
 *     parameters and small scratch.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream). All
 *     work is stream-ordered; functions documented as "synchronous" block the host until
 *     their result is in host memory.
 *   - a context is bound to one GPU and is not thread-safe; use one context per GPU.
 */
#ifndef REINFOCUS_B200_H
#define REINFOCUS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rf_ctx rf_ctx;

typedef enum {
    RF_OK = 0,
    RF_ERR_INVALID = -1, /* bad argument / call order (maps to AssertionError in Python) */
    RF_ERR_CUDA = -2,    /* CUDA runtime failure (maps to RuntimeError) */
    RF_ERR_NOMEM = -3,
    RF_ERR_NO_SCENE = -4 /* render before update_targets/update_focus_planes
                            (reference graphics/device_data.py:43 AssertionError) */
} rf_status;

/* One xoroshiro128+ state, layout-identical to numba.cuda.random.xoroshiro128p_dtype. */
typedef struct {
    uint64_t s0;
    uint64_t s1;
} rf_rng_state;

/* -------------------------------------------------------------------------------------
 * Context
 * ----------------------------------------------------------------------------------- */

/* Creates a context on CUDA device `device`. Replaces the implicit numba context that
 * reference graphics/render.py:127-145 (FastRenderer.__init__) relies on. */
int rf_create(rf_ctx **out, int device);
int rf_destroy(rf_ctx *ctx);
const char *rf_last_error(const rf_ctx *ctx);
const char *rf_last_global_error(void);
/* ABI version of this library (bumped on any signature change). */
int rf_abi_version(void);
/* sizeof() of the structs passed across this boundary, so a binding can check its layout. */
enum { RF_SIZEOF_SCENE_PACKING = 0, RF_SIZEOF_ENV_CONFIG = 1 };
int rf_sizeof(int which);
/* Device facts the host side sizes launches and rooflines with. */
int rf_device_info(const rf_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor,
                   int *clock_khz);
/* Number of kernels this context has launched since creation (bench "gpu_launches"). */
int64_t rf_launch_count(const rf_ctx *ctx);

/* -------------------------------------------------------------------------------------
 * RNG states. Replaces reference graphics/random.py:8-18 (make_random_states ->
 * numba.cuda.random.create_xoroshiro128p_states, a sequential CPU jump chain + H2D) and
 * the cache policy of reference graphics/render.py:248-257 (_make_random_states).
 * ----------------------------------------------------------------------------------- */

/* Guarantees the context holds at least n_states states. If it holds fewer, ALL states
 * are re-created from `seed` at the new size (reference render.py:256-257 re-seeds with 0
 * on growth); otherwise the existing, already-advanced states are kept. */
int rf_rng_ensure(rf_ctx *ctx, int64_t n_states, uint64_t seed, void *stream);
/* Drops the cached states (next rf_rng_ensure re-creates them). */
int rf_rng_reset(rf_ctx *ctx);
int64_t rf_rng_count(const rf_ctx *ctx);
/* Synchronous copies of states [first, first+n) to / from host memory (parity tests,
 * checkpointing of env state - the reference cannot serialise its states). */
int rf_rng_export(rf_ctx *ctx, rf_rng_state *h_dst, int64_t first, int64_t n);
int rf_rng_import(rf_ctx *ctx, const rf_rng_state *h_src, int64_t first, int64_t n);
/* Fills a caller-owned device buffer with create_xoroshiro128p_states(n, seed): state i
 * is state 0 jumped i * 2**64 steps, computed in parallel by GF(2) matrix doubling. */
int rf_rng_init_device(rf_ctx *ctx, rf_rng_state *d_states, int64_t n, uint64_t seed,
                       void *stream);
/* xoroshiro128p_uniform_float32 draws: out[i*draws + k] = k-th draw of state i (advances
 * the caller's states). Test / API-parity helper for reference graphics/random.py:21-33. */
int rf_rng_uniform_device(rf_ctx *ctx, rf_rng_state *d_states, int64_t n, int draws,
                          float *d_out, void *stream);

/* -------------------------------------------------------------------------------------
 * Scene parameters. Replaces the device arrays built by reference
 * graphics/world.py:100-123 (FastWorlds._make_device_data: float32 [n,2] = half side, z)
 * and graphics/camera.py:132-179 (FastCameras._make_device_data: float32 [n,3,3] = lower
 * left, horizontal, vertical, plus static origin/u/v and the float64 lens radius).
 * The host side packs the numbers (bit-equal to the reference's NumPy arithmetic); these
 * calls only move them. h_world / h_cam_dyn may be NULL to leave that part unchanged.
 * ----------------------------------------------------------------------------------- */
int rf_set_world(rf_ctx *ctx, int n, const float *h_world, void *stream);
int rf_set_cameras(rf_ctx *ctx, int n, const float *h_cam_dyn, const float origin[3],
                   const float u[3], const float v[3], double lens_radius, void *stream);
/* Number of envs of the last rf_set_world (reference render.py:174 sizes the batch by
 * len(self._worlds)); 0 before any update. */
int rf_scene_envs(const rf_ctx *ctx);

/* -------------------------------------------------------------------------------------
 * Tracer. Replaces the kernel launch of reference graphics/render.py:165-246
 * (FastRenderer.render / FastRenderer._device_render) for the first `n` envs of the
 * scene: spp samples per pixel, pixel (e, y, x) uses RNG state e*H*W + y*W + x and leaves
 * it advanced. Outputs (either may be NULL, not both):
 *   d_rgb  uint8 [n, H, W, 3]  exactly the reference's frames;
 *   d_gray uint8 [n, H, W]     cv2.cvtColor(RGB2GRAY) of those frames, produced in
 *                              registers without the RGB round trip.
 * Requires rf_rng_ensure(n*H*W) (called internally with seed 0 if needed).
 * ----------------------------------------------------------------------------------- */
int rf_render(rf_ctx *ctx, int n, int H, int W, int spp, uint8_t *d_rgb, uint8_t *d_gray,
              void *stream);

/* -------------------------------------------------------------------------------------
 * General-scene tracer. Replaces reference graphics/render.py:88-119 (render) and :31-85
 * (device_render): spheres and z-aligned rectangles with per-shape checker frequency, up to
 * 50 bounces, one camera per env, fresh RNG states from `seed` on every call (the reference
 * seeds 0, render.py:115). Host inputs, laid out as the reference's Worlds / Cameras build
 * them (world.py:27-82, camera.py:59-91):
 *   h_shape_params float32 [n, max_shapes, 7]  sphere: cx cy cz r fx fy 0
 *                                              rectangle: xmin xmax ymin ymax z fx fy
 *   h_shape_types  int32   [n, max_shapes]     0 sphere, 1 rectangle
 *   h_env_sizes    int32   [n]                 shapes in use per env
 *   h_cameras      float64 [n, 19]             lower-left, horizontal, vertical, origin, u, v
 *                                              (3 each) and the lens radius
 * Output d_rgb uint8 [n, H, W, 3]. Synchronous.
 * ----------------------------------------------------------------------------------- */
int rf_render_generic(rf_ctx *ctx, int n, int H, int W, int spp, int max_shapes,
                      const float *h_shape_params, const int *h_shape_types,
                      const int *h_env_sizes, const double *h_cameras, uint64_t seed,
                      uint8_t *d_rgb, void *stream);

/* -------------------------------------------------------------------------------------
 * Focus measure. Replaces reference vision.py:11-39 (focus_value / focus_values:
 * cv2.cvtColor -> cv2.medianBlur(3) -> cv2.Laplacian(CV_8U) -> ndarray.var()).
 * d_img is uint8 [n, H, W, channels] with channels 1 (gray) or 3 (RGB);
 * d_out float64 [n].
 * ----------------------------------------------------------------------------------- */
int rf_focus(rf_ctx *ctx, int n, int H, int W, const uint8_t *d_img, int channels,
             double *d_out, void *stream);
/* Debug/parity variant that also writes the intermediate planes (uint8 [n, H, W] each,
 * may be NULL). */
int rf_focus_planes(rf_ctx *ctx, int n, int H, int W, const uint8_t *d_img, int channels,
                    double *d_out, uint8_t *d_median, uint8_t *d_laplacian, void *stream);

/* -------------------------------------------------------------------------------------
 * One hot-path step = what reference environments/state_observer.py:359-383
 * (FocusObserver.observe) does per call: set targets/focus planes, render, focus values.
 * Host buffers in, host focus values out; the H2D/D2H copies, both kernels and the final
 * synchronisation are all inside the call ("synchronous"). Frames never leave the GPU.
 *   h_world float32 [n,2], h_cam_dyn float32 [n,9] (as rf_set_world / rf_set_cameras),
 *   h_focus float64 [n].
 * rf_step_device is the same with the focus values left on the device (no sync).
 * ----------------------------------------------------------------------------------- */
int rf_step_host(rf_ctx *ctx, int n, int H, int spp, const float *h_world,
                 const float *h_cam_dyn, double *h_focus, void *stream);
int rf_step_device(rf_ctx *ctx, int n, int H, int spp, double *d_focus, void *stream);

/* -------------------------------------------------------------------------------------
 * Scene parameters from device-resident positions. Replaces the host packing of reference
 * graphics/world.py:107-123 (FastWorlds._make_device_data) and graphics/camera.py:132-179
 * (FastCameras._make_device_data) plus their uploads: env i takes its target position from
 * d_targets[i * stride] and its focus-plane distance from d_planes[i * stride] (so an
 * [n, 2] state tensor is passed as base, base + 1, stride 2). The float32 operations are
 * the reference's, in its order; the constants are the float32 values the host classes
 * hold: world_tan = tan(radians(r_size / 2)), half_* = tan(vfov / 2) (* aspect),
 * full_* = float32(2 * half), and the camera frame origin / u / v / w.
 * ----------------------------------------------------------------------------------- */
typedef struct {
    float world_tan;
    float half_width, half_height, full_width, full_height;
    float origin[3], u[3], v[3], w[3];
    double lens_radius;
} rf_scene_packing;
int rf_set_scene_device(rf_ctx *ctx, int n, const float *d_targets, const float *d_planes,
                        int stride, const rf_scene_packing *packing, void *stream);
/* rf_step_host for callers that hold the positions themselves, not packed parameters: host
 * targets / focus planes in (float32 [n] each), host float64 focus values out; the packing of
 * FastWorlds / FastCameras._make_device_data runs on the GPU (rf_set_scene_device). One call
 * = one FocusObserver.observe (reference environments/state_observer.py:377-383).
 * Synchronous. */
int rf_step_positions_host(rf_ctx *ctx, int n, int H, int spp, const float *h_targets, const float *h_planes,
                           const rf_scene_packing *packing, double *h_focus, void *stream);

/* -------------------------------------------------------------------------------------
 * Device-resident vector env. Replaces reference environments/vector_environment.py:75-164
 * (VectorEnvironment.reset / step) for the strategy compositions of the example envs
 * (reference examples/custom_environments.py): state [target, focus plane];
 *   transformer  RF_ENV_DISCRETE_MOVE     DiscreteMoveTransformer (state_transformer.py:222-266)
 *                RF_ENV_CONTINUOUS_JUMP   ContinuousJumpTransformer (:88-137)
 *                RF_ENV_CONTINUOUS_MOVE   ContinuousMoveTransformer (:140-192)
 *                RF_ENV_DISCRETE_JUMP     DiscreteJumpTransformer (:195-219)
 *   ender        any tree of TimeLimitEnder, DivergingEnder, OnTargetEnder, StoppedEnder,
 *                EndlessEnder combined with & and | (episode_ender.py:112-656), given as a
 *                postfix program of rf_env_ender nodes
 *   observer     any tree of DeltaObservers and NormalizedObservers over
 *                IndexedElementObservers and one FocusObserver (state_observer.py:103-517),
 *                given as a postfix program of rf_env_observer nodes
 *   rewarder     any tree of DeltaRewarder, DistanceRewarder, ObservationRewarder,
 *                OnTargetRewarder, StoppedRewarder combined with + and *
 *                (episode_rewarder.py:86-429), as a postfix program of rf_env_reward nodes;
 *                NumPy's result types are followed (float32 until a float64 operand joins)
 *   initializer  RangedInitializer (up to 4 ranges per element), numpy PCG64DXSM generator
 *                (state_initializer.py:30-71)
 * States, observations, rewards and the generator live on the GPU; a step is two small
 * kernels around the render + focus launches of the context (and of the k restarted envs,
 * rendered as batch positions 0..k-1 like the reference's partial reset). Sequences are
 * bit-identical to the host classes driven by the same generator.
 * ----------------------------------------------------------------------------------- */
typedef struct rf_env rf_env;
enum { RF_ENV_DISCRETE_MOVE = 0, RF_ENV_CONTINUOUS_JUMP = 1, RF_ENV_CONTINUOUS_MOVE = 2, RF_ENV_DISCRETE_JUMP = 3 };
enum { RF_ENV_ENDER_TIME_LIMIT = 0, RF_ENV_ENDER_DIVERGING = 1, RF_ENV_ENDER_ON_TARGET = 2,
       RF_ENV_ENDER_STOPPED = 3, RF_ENV_ENDER_ENDLESS = 4, RF_ENV_ENDER_AND = 5, RF_ENV_ENDER_OR = 6 };
enum { RF_ENV_REWARD_DELTA = 0, RF_ENV_REWARD_DISTANCE = 1, RF_ENV_REWARD_OBSERVATION = 2,
       RF_ENV_REWARD_ON_TARGET = 3, RF_ENV_REWARD_STOPPED = 4, RF_ENV_REWARD_ADD = 5, RF_ENV_REWARD_MUL = 6 };
/* One node of the ender program. i0 / i1: the state elements it looks at; steps: max_steps
 * (time limit) or early_end_steps (<= 15 for STOPPED); value: threshold / radius / span. */
typedef struct {
    int kind, i0, i1, steps;
    float value;
} rf_env_ender;
/* One node of the rewarder program. DELTA: i0, f0 = reward, f1 = scale. DISTANCE: i0, i1,
 * f0 = span, f1 = float32(high - low), d0 = low. OBSERVATION: i0 = observation column.
 * ON_TARGET: i0, i1, f0 = span, d0 = off, d1 = on - off. STOPPED: i0, f0 = threshold,
 * d0 = reward. */
typedef struct {
    int kind, i0, i1;
    float f0, f1;
    double d0, d1;
} rf_env_reward;
/* One node of the observer program, children before parents. ELEMENT: arg = the state
 * element an IndexedElementObserver shows. FOCUS: the FocusObserver (exactly one per
 * program). DELTA / NORMALIZED: arg = number of child observers (their observations side by
 * side, as WrapperObserver stacks them); DELTA: flag = include_original; NORMALIZED: offset =
 * first of its columns in obs_mid / obs_scale (one (_mid, _scale) pair per column it sees). */
enum { RF_ENV_OBS_ELEMENT = 0, RF_ENV_OBS_FOCUS = 1, RF_ENV_OBS_DELTA = 2, RF_ENV_OBS_NORMALIZED = 3 };
typedef struct {
    int kind, arg, flag, offset;
} rf_env_observer;
enum { RF_ENV_MAX_NODES = 24, RF_ENV_MAX_OBS_NODES = 16, RF_ENV_MAX_OBS_DIM = 16, RF_ENV_MAX_OBS_VALUES = 32 };
enum { RF_ENV_ACTIONS_INT32 = 0, RF_ENV_ACTIONS_INT64 = 1, RF_ENV_ACTIONS_FLOAT32 = 2 };
typedef struct {
    int num_envs, frame_height, samples_per_pixel;
    int transformer;
    int n_moves;             /* <= 32 */
    double moves[32];        /* discrete action set, float64 as the reference keeps it */
    float limits[2];         /* clip range of the discrete move / ends of the jump */
    float jump_span;         /* float32(limits[1] - limits[0]) */
    float jump_threshold;    /* jumps / continuous moves shorter than this are ignored */
    float move_speed;        /* RF_ENV_CONTINUOUS_MOVE: distance of action 1.0 */
    float jumps[32];         /* RF_ENV_DISCRETE_JUMP: positions, float32 as the reference keeps them */
    int n_enders, n_rewards; /* program lengths, 1..RF_ENV_MAX_NODES each */
    rf_env_ender enders[RF_ENV_MAX_NODES];
    rf_env_reward rewards[RF_ENV_MAX_NODES];
    /* observer program (1..RF_ENV_MAX_OBS_NODES nodes; the root is the last node). An
     * observation has at most RF_ENV_MAX_OBS_DIM columns and no more than
     * RF_ENV_MAX_OBS_VALUES values are alive at any point of the program. obs_mid /
     * obs_scale: the NormalizedObservers' _mid / _scale rows, at their nodes' offsets. */
    int n_observers;
    rf_env_observer observers[RF_ENV_MAX_OBS_NODES];
    float obs_mid[RF_ENV_MAX_OBS_VALUES], obs_scale[RF_ENV_MAX_OBS_VALUES];
    /* RangedInitializer: per state element 1..4 (low, high) ranges; with several, the range
     * is picked like Generator.choice does before Generator.uniform draws inside it */
    int init_options[2];
    double init_low[2][4], init_high[2][4];
    rf_scene_packing packing;
} rf_env_config;
int rf_env_create(rf_ctx *ctx, const rf_env_config *config, rf_env **out);
int rf_env_destroy(rf_env *env);
/* numpy.random.PCG64DXSM().state: 128-bit state and increment as (high, low) words, plus
 * the buffered 32-bit half (has_uint32, uinteger) that Generator.choice draws from. */
int rf_env_set_generator(rf_env *env, const uint64_t state[2], const uint64_t inc[2], uint32_t has_uint32,
                         uint32_t uinteger);
int rf_env_get_generator(rf_env *env, uint64_t state[2], uint64_t inc[2], uint32_t *has_uint32,
                         uint32_t *uinteger); /* synchronous */
/* Columns of an observation (rf_env_reset / rf_env_step write float32 [n, rf_env_obs_dim]). */
int rf_env_obs_dim(const rf_env *env);
/* reset(): every env starts an episode; d_obs float32 [n, obs_dim]. */
int rf_env_reset(rf_env *env, float *d_obs, void *stream);
/* step(): d_actions [n] of the given kind; d_obs float32 [n, obs_dim], d_rewards float64 [n],
 * d_truncated uint8 [n] (nothing ever terminates, as in the reference); *h_resets, if not
 * NULL, receives the number of envs that restarted. The host waits only for that count,
 * which is known before the main render starts; the outputs are stream-ordered. Drive an
 * env from one stream: its state lives in device memory and is ordered only by that stream. */
int rf_env_step(rf_env *env, const void *d_actions, int action_kind, float *d_obs,
                double *d_rewards, uint8_t *d_truncated, int *h_resets, void *stream);
/* Synchronous copies of the env's episode state to / from host memory (parity tests,
 * checkpoint / resume; the reference cannot serialise an env). Together with the generator
 * (rf_env_get/set_generator) and the renderer's RNG states (rf_rng_export/import) this is
 * everything a resumed run needs to continue bit-identically. h_states float32 [n, 2],
 * h_old_obs float32 [n, rf_env_delta_width(env)] (the DeltaObservers' previous values),
 * h_node_state uint32 [rf_env_node_rows(env), n]
 * (the enders' counters / windows and the rewarders' previous values, raw bits). Export
 * skips NULL pointers; import needs all of them and stands in for a reset. */
int rf_env_node_rows(const rf_env *env);
int rf_env_delta_width(const rf_env *env);
int rf_env_export(rf_env *env, float *h_states, float *h_old_obs, uint32_t *h_node_state);
int rf_env_import(rf_env *env, const float *h_states, const float *h_old_obs, const uint32_t *h_node_state);

/* -------------------------------------------------------------------------------------
 * Self-checks and measurement helpers (used by tests/ and bench.py).
 * ----------------------------------------------------------------------------------- */
/* Exhaustive device-side checks of the arithmetic shortcuts the tracer takes; *mismatches
 * receives the number of inputs on which the shortcut differs from the literal reference
 * expression (must be 0):
 *   RF_SELFTEST_CHECKER     table-based checkerboard cell vs sin(32*pi*u) in float64, for
 *                           every float32 u in [0, 1];
 *   RF_SELFTEST_PIXEL_DIV   hoisted-reciprocal float64 division vs __ddiv_rn for
 *                           float32((x + U) / arg), every x in [0, arg), every float32 U in
 *                           [0, 1] (arg = frame width or height);
 *   RF_SELFTEST_INV_LENGTH  branch-free 1/sqrt-length vs __frcp_rn(__fsqrt_rn()) for every
 *                           float32 in [2^-60, 2^60];
 *   RF_SELFTEST_CONST_DIV   hoisted-reciprocal float32 division vs __fdiv_rn for every
 *                           numerator in [0, c], over `arg` divisors c spread over [0.5, 8);
 *   RF_SELFTEST_CHECKER_PAIR  the multi-pixel tracer's FMA-pipe checker parity (RZ / RU fused
 *                           multiply-adds) vs the table-based cell, every float32 u in [0, 1]. */
enum {
    RF_SELFTEST_CHECKER = 0,
    RF_SELFTEST_PIXEL_DIV = 1,
    RF_SELFTEST_INV_LENGTH = 2,
    RF_SELFTEST_CONST_DIV = 3,
    RF_SELFTEST_CHECKER_PAIR = 4
};
int rf_selftest(rf_ctx *ctx, int which, int arg, int64_t *mismatches, void *stream);
/* Options. RF_OPT_FORCE_GENERIC = 1 makes rf_render use the literal any-camera kernel and
 * rf_focus the general staged kernel even when the specialised ones apply (A/B parity
 * tests). RF_OPT_TRACE_CONTEXTS = -1 (default), 0 or 2..8: pixels per thread of the
 * default-camera tracer (0 = the one-pixel-per-thread kernel, -1 = by batch size: 0 for one
 * or two 300 x 300 envs, then 4 and 8 pixels per thread in blocks of 64 to 256 threads);
 * every setting produces the same bytes. */
enum { RF_OPT_FORCE_GENERIC = 0, RF_OPT_TRACE_CONTEXTS = 1 };
int rf_set_option(rf_ctx *ctx, int option, int value);
/* Introspection. RF_INFO_LAST_TRACE_KERNEL: 1 if the last rf_render used the specialised
 * one-pixel-per-thread kernel, 2..8 for the multi-pixel kernel with that many pixels
 * per thread, 0 for the generic one, -1 before any render. RF_INFO_LAST_FOCUS_KERNEL: 1 if the
 * last rf_focus used the packed kernel, 0 for the staged one. */
enum { RF_INFO_LAST_TRACE_KERNEL = 0, RF_INFO_LAST_FOCUS_KERNEL = 1 };
int rf_get_info(const rf_ctx *ctx, int what);
/* Measures the FP32 FFMA peak of this GPU (dependent-chain-free FFMA loop on all SMs);
 * returns TFLOP/s counting 2 flop per FFMA. */
int rf_measure_fp32_peak(rf_ctx *ctx, double *tflops, double *sm_clock_mhz_seen);

#ifdef __cplusplus
}
#endif

#endif
