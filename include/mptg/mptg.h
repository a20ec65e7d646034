/* mptg.h -- C ABI of the B200-native planning hot path (libmptg.so).
 *
 * Drop-in boundary for UNC-Robotics/mpt's data-parallel hot path.  Each entry point names the
 * reference interface it replaces (paths relative to the reference root).  No torch / C++ types in
 * any signature: plain pointers, sizes and opaque handles.  Every function returns an int status
 * (MPTG_OK == 0, negative = error); mptg_last_error() gives the text.  There is NO CPU fallback: if
 * the CUDA device, the kernel image or a launch is unavailable the call fails with MPTG_ERR_CUDA.
 *
 * Conventions
 *  - A "state" is the concatenation of the scalars of the space's parts, in part order:
 *      LP(dim)  : dim scalars                     (Eigen::Matrix<S,dim,1>, src/mpt/lp_space.hpp:44-54)
 *      SO2(dim) : dim angles                      (src/mpt/so2_space.hpp:44-65)
 *      SO3      : quaternion coeffs (x,y,z,w)     (Eigen::Quaternion::coeffs(), src/mpt/so3_space.hpp:52-55)
 *    e.g. SE(3) = {SO3 weight 50, LP(p=2,dim=3) weight 1}: 7 scalars qx qy qz qw tx ty tz
 *    (rotation is tuple element 0, src/mpt/se3_space.hpp:53-79).
 *  - Host buffers are array-of-states (AoS), row-major, `scalar` bytes per element (4 or 8).
 *  - Node identity is the dense uint32 index in insertion order (the host keeps index -> Node*).
 *  - Functions ending in _dev take DEVICE pointers valid on the context's device and are enqueued
 *    on the context stream without synchronising; the plain forms take HOST pointers, do the
 *    host<->device copies on the context stream and return after the results are in host memory.
 *  - A handle is single-owner: one host thread drives one context (one GPU, one stream).
 */
#ifndef MPTG_H
#define MPTG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPTG_ABI_VERSION 2 /* 2: near_contact_out on mptg_valid_batch / mptg_link_batch; mptg_comm_*, mptg_knn_query_sharded */
#define MPTG_MAX_PARTS 8
#define MPTG_MAX_SCALARS 64 /* max scalars per state */
#define MPTG_MAX_K 128      /* max neighbours per query */
#define MPTG_NO_INDEX 0xFFFFFFFFu

enum mptg_status {
    MPTG_OK = 0,
    MPTG_ERR_BAD_ARG = -1,
    MPTG_ERR_OOM = -2,
    MPTG_ERR_CUDA = -3,
    MPTG_ERR_NCCL = -4,
    MPTG_ERR_CAPACITY = -5,
    MPTG_ERR_UNSUPPORTED = -6
};

enum mptg_part_kind { MPTG_PART_LP = 1, MPTG_PART_SO2 = 2, MPTG_PART_SO3 = 3 };
enum mptg_scalar { MPTG_F32 = 4, MPTG_F64 = 8 };
enum mptg_knn_strategy { MPTG_KNN_AUTO = 0, MPTG_KNN_BRUTE = 1, MPTG_KNN_BVH = 2 };
enum mptg_geom_kind { MPTG_GEOM_GRID = 1, MPTG_GEOM_SHAPES = 2, MPTG_GEOM_LINKARM = 3, MPTG_GEOM_MESH = 4, MPTG_GEOM_NAOCUP = 5 };

/* One factor of a Cartesian product space.  Replaces the compile-time metric tags
 * LP<p>, SO2<p>, SO3, Scaled<M,ratio>, Cartesian<...> (src/mpt/impl/metrics.hpp:40-42,
 * src/mpt/{lp,so2,so3,scaled,cartesian}_space.hpp).  distance = sum_i weight_i * d_i. */
typedef struct mptg_space_part {
    int32_t kind;  /* mptg_part_kind */
    int32_t p;     /* LP/SO2 norm: 1, 2, or 0 for infinity */
    int32_t dim;   /* LP/SO2: number of scalars; SO3: ignored (4 scalars) */
    int32_t _pad;
    double weight; /* Scaled<> ratio; 1.0 when unscaled.  Applied as d*weight in `scalar` precision */
} mptg_space_part;

typedef struct mptg_space_desc {
    int32_t n_parts;
    int32_t scalar; /* mptg_scalar */
    mptg_space_part part[MPTG_MAX_PARTS];
} mptg_space_desc;

typedef struct mptg_ctx mptg_ctx;
typedef struct mptg_knn mptg_knn;
typedef struct mptg_geom mptg_geom;

/* ------------------------------------------------------------------ context */
int mptg_abi_version(void);
/* device < 0 selects the current CUDA device.  Creates the context stream. */
int mptg_ctx_create(int device, mptg_ctx** out);
int mptg_ctx_destroy(mptg_ctx* ctx);
int mptg_sync(mptg_ctx* ctx);
/* Text of the last error on this context (ctx may be NULL for creation failures). Never NULL. */
const char* mptg_last_error(const mptg_ctx* ctx);
/* cudaStream_t of the context, as an opaque pointer (so callers can order their own work). */
void* mptg_ctx_stream(mptg_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t mptg_ctx_launch_count(const mptg_ctx* ctx);
/* Streaming multiprocessors of the context's device (grids are sized in multiples of it; bench.py's issue-slot roofline). */
int mptg_ctx_sm_count(const mptg_ctx* ctx);
/* Measured FP32 rate of this GPU: a register-only FFMA kernel (8 independent chains per thread, every SM full),
 * best of two timed launches of ~4 ms, in TFLOP/s (2 flop per FFMA).  The yardstick bench.py's fp32 rooflines use
 * ("FP32 peak is not in MEASURED_PEAKS.json; measure it with an FFMA microbenchmark"). */
int mptg_probe_fp32_tflops(mptg_ctx* ctx, double* tflops_out);
/* Scalars per state for a space (sum of part sizes); < 0 on a malformed descriptor. */
int mptg_space_scalars(const mptg_space_desc* space);
/* space.dimensions() of the reference (LP: dim, SO2: dim, SO3: 3, sum over parts) used by
 * src/mpt/impl/rrg_rewire_neighbors.hpp:60 and src/mpt/impl/pprm/pprm.hpp:146. */
int mptg_space_dimensions(const mptg_space_desc* space);

/* ------------------------------------------------------ metric (a4, a5) */
/* Space::distance(a,b) for n state pairs.  Replaces nigh::metric::Space<T,M>::distance as used at
 * src/mpt/discrete_motion_validator.hpp:78, impl/prrt_star/prrt_star.hpp:535, goal_state.hpp:65. */
int mptg_distance_batch(mptg_ctx* ctx, const mptg_space_desc* space, const void* a, const void* b,
                        uint32_t n, void* dist_out);
/* interpolate(space,a,b,t) for n triples; t has n scalars.  Replaces the overloads at
 * src/mpt/lp_space.hpp:44-54, so2_space.hpp:44-65, so3_space.hpp:44-83, scaled_space.hpp:44-51,
 * cartesian_space.hpp:44-67. */
int mptg_interpolate_batch(mptg_ctx* ctx, const mptg_space_desc* space, const void* a, const void* b,
                           const void* t, uint32_t n, void* out);

/* ------------------------------------------------------------ kNN (a1-a3) */
/* nigh::Nigh<Node*,Space,NodeKey,Concurrency,Strategy> nn(space)  (impl/prrt/prrt.hpp:121-122). */
int mptg_knn_create(mptg_ctx* ctx, const mptg_space_desc* space, uint32_t capacity, mptg_knn** out);
int mptg_knn_destroy(mptg_knn* knn);
int mptg_knn_set_strategy(mptg_knn* knn, int strategy /* mptg_knn_strategy */);
/* Multi-GPU sharding: reported index = local_index * mul + add (default 1, 0).  With tree points
 * dealt round-robin to G ranks (global index g lives on rank g % G at local slot g / G) rank r
 * sets mul = G, add = r and every result carries global indices. */
int mptg_knn_set_index_map(mptg_knn* knn, uint32_t mul, uint32_t add);
/* nn.insert(node) for a batch (impl/prrt/prrt.hpp:186,447; prrt_star.hpp:278,619; pprm.hpp:337).
 * Indices first .. first+count-1 are assigned in order. */
int mptg_knn_insert(mptg_knn* knn, const void* states, uint32_t count, uint32_t* first_index_out);
int mptg_knn_insert_dev(mptg_knn* knn, const void* states_dev, uint32_t count, uint32_t* first_index_out);
/* nn.size() */
uint32_t mptg_knn_size(const mptg_knn* knn);
/* Read back `count` stored states starting at index `first` (AoS, host). */
int mptg_knn_get_states(mptg_knn* knn, uint32_t first, uint32_t count, void* states_out);
/* nn.nearest(q) (k == 1) and nn.nearest(nbh, q, k, r) (impl/prrt/prrt.hpp:406-409,
 * prrt_star.hpp:505-508, rrg_rewire_neighbors.hpp:65-67,125-128, pprm.hpp:304) for Q queries.
 * Result row q holds count_out[q] <= k neighbours, ascending by (distance, index); unused slots are
 * MPTG_NO_INDEX / +inf.  radius < 0 or +inf means unbounded; otherwise only distance <= radius
 * is kept (if more than k qualify the k nearest are returned and count_out[q] == k).
 * dist_out has Q*k scalars of the space's scalar type.  count_out may be NULL. */
int mptg_knn_query(mptg_knn* knn, const void* queries, uint32_t Q, uint32_t k, double radius,
                   uint32_t* idx_out, void* dist_out, uint32_t* count_out);
int mptg_knn_query_dev(mptg_knn* knn, const void* queries_dev, uint32_t Q, uint32_t k, double radius,
                       uint32_t* idx_out_dev, void* dist_out_dev, uint32_t* count_out_dev);
/* Build / refresh the spatial index now (otherwise done lazily by the AUTO strategy). */
int mptg_knn_build_index(mptg_knn* knn);
/* Counters of the last query call: [0]=distance evaluations, [1]=index nodes visited, [2]=indexed
 * points, [3]=strategy used (mptg_knn_strategy). */
int mptg_knn_last_stats(mptg_knn* knn, uint64_t stats_out[4]);
/* Multi-GPU merge step: `parts` candidate lists per query (each k wide, ascending, global indices),
 * laid out [parts][Q][k] in device memory, merged to the k best by (distance, index).
 * Used after an all-gather of per-shard results (tree points sharded across GPUs). */
int mptg_knn_merge_dev(mptg_ctx* ctx, int scalar, uint32_t parts, uint32_t Q, uint32_t k,
                       const uint32_t* idx_in_dev, const void* dist_in_dev, uint32_t* idx_out_dev,
                       void* dist_out_dev, uint32_t* count_out_dev);

/* ------------------------------------------------ tree-sharded kNN across GPUs
 * The reference has one nigh::Nigh structure per planner in host memory; a tree that is spread over the GPUs of a box
 * keeps its interface (insert, nearest) and adds a communicator.  One process (and one mptg_ctx) per GPU.  Every rank
 * stores a SPATIAL shard of the points -- which rank stores a point is the caller's choice (e.g. the sub-box of the
 * sampling bounds it falls in); pruning works to the extent the shards are spatially compact, results are exact for any
 * assignment -- and all ranks call mptg_knn_query_sharded collectively with the SAME query wave.  Rank r receives the
 * results of its slice of the wave (mptg_comm_slice): rows ascending by (distance, global index), identical to what a
 * single structure holding all points returns.  Collectives are NCCL (all-gather of the shards' top-level boxes at
 * mptg_knn_shard_sync; per wave an all-reduce(min) of the per-query distance bounds and a grouped send/recv of the
 * candidate rows), resolved at run time
 * from libnccl.so.2; any NCCL failure returns MPTG_ERR_NCCL. */
#define MPTG_UNIQUE_ID_BYTES 128
typedef struct mptg_comm mptg_comm;
/* ncclGetUniqueId: called by one rank, the 128 bytes are handed to the others by the caller (MPI, torch.distributed,
 * a file ...) and passed to mptg_comm_init by every rank. */
int mptg_comm_unique_id(void* id_out);
int mptg_comm_init(mptg_ctx* ctx, const void* unique_id, int rank, int world, mptg_comm** out);
int mptg_comm_destroy(mptg_comm* comm);
int mptg_comm_rank(const mptg_comm* comm);
int mptg_comm_world(const mptg_comm* comm);
/* The contiguous slice [first, first + count) of n units (queries of a wave) owned by this rank. */
int mptg_comm_slice(const mptg_comm* comm, uint32_t n, uint32_t* first_out, uint32_t* count_out);
/* nn.insert for a shard: `ids[i]` is the index results report for state i (the node's global, insertion-order index in
 * the planner's graph).  A structure takes either plain inserts or inserts with ids, not both. */
int mptg_knn_insert_ids(mptg_knn* knn, const void* states, const uint32_t* ids, uint32_t count);
/* Collective, after inserting and before searching: every rank indexes what its shard stores and the bounding boxes of
 * the children of every shard's top node are exchanged (a few KB), so that each rank can bound the distance of a query
 * to every shard by itself.  mptg_knn_query_sharded fails with MPTG_ERR_BAD_ARG if the shard has changed since. */
int mptg_knn_shard_sync(mptg_comm* comm, mptg_knn* shard);
/* nn.nearest over the union of all ranks' shards, for the Q queries of the wave (the same on every rank).  Outputs hold
 * the rows of this rank's slice only: count * k entries.  Semantics of k / radius / padding as mptg_knn_query. */
int mptg_knn_query_sharded(mptg_comm* comm, mptg_knn* shard, const void* queries, uint32_t Q, uint32_t k, double radius,
                           uint32_t* idx_out, void* dist_out, uint32_t* count_out);
int mptg_knn_query_sharded_dev(mptg_comm* comm, mptg_knn* shard, const void* queries_dev, uint32_t Q, uint32_t k, double radius,
                               uint32_t* idx_out_dev, void* dist_out_dev, uint32_t* count_out_dev);

/* ------------------------------------------------ scenario geometry (a7-a10) */
/* Occupancy grid, 1 byte per cell (non-zero = obstacle), row-major width*height.
 * Replaces PNG2dScenario's std::vector<bool> isObstacle_ (demo/png_2d_scenario.hpp:86,104-110).
 * States are LP(2): (x, y) in pixels.  Cells whose linear index falls outside [0, w*h) -- the
 * reference reads out of bounds there -- are treated as obstacles. */
int mptg_grid_create(mptg_ctx* ctx, int scalar, int32_t width, int32_t height, const uint8_t* occupancy,
                     mptg_geom** out);
/* Balls (centres in `dim`-D, point and closed-form segment tests) and 2-D rectangles (bisected to
 * 1 unit).  Replaces shape::Circle / shape::Rect (demo/shape_hierarchy.hpp:168-273), the scenario
 * loops of demo/holonomic_2d_point_scenario.hpp:95-113 and the sphere scenario of
 * test/planner_integration_test.hpp:128-150.  centres: n_balls*dim doubles; radii: n_balls doubles;
 * rects: n_rects*4 doubles (x0,y0,x1,y1), only with dim == 2. */
int mptg_shapes_create(mptg_ctx* ctx, int scalar, int32_t dim, int32_t n_balls, const double* centres,
                       const double* radii, int32_t n_rects, const double* rects, mptg_geom** out);
/* Planar N-link arm among circles.  Replaces LinkManipulatorScenario::valid/link/bisectLink
 * (demo/link_manipulator_scenario.hpp:99-138).  States are LP(n_links) joint angles. */
int mptg_linkarm_create(mptg_ctx* ctx, int scalar, int32_t n_links, const double* lengths,
                        double link_radius, int32_t n_circles, const double* cx_cy_r, mptg_geom** out);
/* Rigid robot mesh vs static environment mesh (triangle soups, 9 floats per triangle).  Replaces
 * fcl::BVHModel<OBBRSS> + fcl::collide in SE3RigidBodyScenario::valid
 * (demo/se3_rigid_body_scenario.hpp:164-204,282-296).  The robot mesh is used as given (the
 * reference recentres it on the vertex mean at load time, :181-193 -- do that before calling).
 * States are SE(3): qx qy qz qw tx ty tz. */
int mptg_mesh_pair_create(mptg_ctx* ctx, int scalar, uint32_t n_tri_robot, const float* robot_tris,
                          uint32_t n_tri_env, const float* env_tris, mptg_geom** out);
/* The Nao humanoid holding a cup and a ball: 10 joint angles (right then left shoulder pitch, shoulder roll, elbow yaw,
 * elbow roll, wrist yaw), forward kinematics of both arms, 209 sphere / capsule pair tests, the cup must stay upright.
 * Replaces NaoCupScenario::valid / link (demo/nao_cup_planning.cpp:146-152), i.e. nao_clear and nao_link with
 * compute / check_collisions (demo/nao_cup/src/naocup.hpp:556-730,795-840) and the primitives of
 * demo/nao_cup/src/collide.hpp:45-115, linear.hpp:125-150.  The robot, the cup and the obstacles are the reference's
 * constants (naocup.hpp:55-80,222-251,426-553).  States are LP(10) joint angles in radians. */
int mptg_naocup_create(mptg_ctx* ctx, int scalar, mptg_geom** out);
/* The reference's start, goal and joint limits (naocup.hpp:254-301), 10 doubles each (the values a float build of the
 * reference uses are these rounded to float).  Any pointer may be NULL. */
int mptg_naocup_configs(int scalar, double* start, double* goal, double* lo, double* hi);
int mptg_geom_destroy(mptg_geom* geom);
int mptg_geom_kind(const mptg_geom* geom);

/* scenario.valid(q) for n states -> ok_out[i] in {0,1}.  (impl/prrt/prrt.hpp:439,
 * prrt_star.hpp:539, pprm.hpp:299)
 * near_contact_out (optional, NULL to skip): 1 where the decision of item i rests on a configuration within the
 * geometry's contact band (mptg_geom_contact_band) of touching -- some triangle pair of a checked state whose largest
 * normalised separating-axis gap lies in (-band, band].  Decisions equal the exact ones outside the band; inside it they
 * are reported here instead of being trusted (SURVEY.md 8d "correctness gates").  For an edge the flag covers the
 * states the call examined: every state of an edge reported valid, the states up to the first collision otherwise.
 * Always 0 for grids, shapes and link arms, whose decisions are bit-identical to the reference's arithmetic. */
int mptg_valid_batch(mptg_geom* geom, const void* states, uint32_t n, uint8_t* ok_out, uint8_t* near_contact_out);
int mptg_valid_batch_dev(mptg_geom* geom, const void* states_dev, uint32_t n, uint8_t* ok_out_dev, uint8_t* near_contact_out_dev);
/* The contact band of a geometry, as an absolute length: 1e-6 x the diagonal of the environment mesh's bounding box
 * (0 for the other geometry kinds). */
int mptg_geom_contact_band(const mptg_geom* geom, double* band_out);
/* scenario.link(a,b) for n edges -> ok_out[i] in {0,1}.  (impl/prrt/prrt.hpp:454-457,
 * prrt_star.hpp:659-662, pprm.hpp:364-366).  Semantics per geometry kind:
 *   GRID    valid(a) && valid(b) && midpoint bisection until |b-a|^2 < 1   (png_2d_scenario.hpp:112-117,152-165)
 *   SHAPES  balls: closed-form point-segment distance; rects: endpoints + bisection (shape_hierarchy.hpp:184-203,228-270)
 *   LINKARM valid(a) && valid(b) && bisection until |a-b|_inf < 0.02        (link_manipulator_scenario.hpp:118-138)
 *   NAOCUP  midpoint bisection until |b-a|_2 < 1 degree; the ends are NOT checked (naocup.hpp:809-840); an edge whose
 *           length is not finite is invalid (the reference's recursion need not terminate on NaN / infinite joints)
 *   MESH    DiscreteMotionValidator with step size `step` over `space`      (discrete_motion_validator.hpp:71-130);
 *           `from` is assumed valid and not checked, exactly as the reference (:72-73).
 * `space` / `step` are only read for MESH (pass NULL / 0 otherwise).
 * near_contact_out: see mptg_valid_batch. */
int mptg_link_batch(mptg_geom* geom, const mptg_space_desc* space, const void* from, const void* to,
                    uint32_t n, double step, uint8_t* ok_out, uint8_t* near_contact_out);
int mptg_link_batch_dev(mptg_geom* geom, const mptg_space_desc* space, const void* from_dev,
                        const void* to_dev, uint32_t n, double step, uint8_t* ok_out_dev, uint8_t* near_contact_out_dev);
/* Counters of the last valid/link call on this geometry: [0]=states checked, [1]=BV pair tests,
 * [2]=primitive (triangle-pair / circle / cell) tests, [3]=work items. */
int mptg_geom_last_stats(mptg_geom* geom, uint64_t stats_out[4]);

/* --------------------------------------------------------- planner stages */
/* Steer: out[i] = d[i] > range ? interpolate(space, near[i], sample[i], range/d[i]) : sample[i]
 * (impl/prrt/prrt.hpp:430-434, impl/prrt_star/prrt_star.hpp:529-536).  If dist_out != NULL it
 * receives distance(near[i], out[i]) recomputed after steering (PRRT* does, :535; PRRT does not). */
int mptg_steer_batch(mptg_ctx* ctx, const mptg_space_desc* space, const void* near, const void* sample,
                     const void* d, uint32_t n, double range, void* out, void* dist_out);

/* --------------------------------------------------------- sampling
 * Uniform samples of a space: LP coordinates uniform in [lo,hi) (src/mpt/uniform_box_sampler.hpp:60-68), SO2
 * coordinates uniform in [-pi,pi) (impl/uniform_sampler_so2.hpp:58-64), SO3 by the reference's three-uniform
 * formula (impl/uniform_sampler_so3.hpp:55-68), compound spaces part by part
 * (impl/uniform_sampler_cartesian.hpp:75-78).  lo / hi hold one double per scalar of the state (entries of SO2 /
 * SO3 parts are ignored).  Sample number g = first + i is a pure function of (seed, g): Philox4x32-10 counter
 * stream, see csrc/plan.cu -- the reference seeds one mt19937_64 per worker from std::random_device, so there is
 * no reference sequence, only the distributions.  Uniform 0 of every sample is reserved for the goal-bias draw. */
int mptg_space_uniforms(const mptg_space_desc* space); /* uniforms consumed per state (SO3: 3, else 1 per scalar) */
int mptg_sample_batch(mptg_ctx* ctx, const mptg_space_desc* space, const double* lo, const double* hi,
                      uint64_t seed, uint64_t first, uint32_t n, void* states_out);
int mptg_sample_batch_dev(mptg_ctx* ctx, const mptg_space_desc* space, const double* lo, const double* hi,
                          uint64_t seed, uint64_t first, uint32_t n, void* states_out_dev);
/* The deterministic half alone: n * mptg_space_uniforms() uniforms in [0,1) (the space's scalar type) -> n
 * states.  With all-zero uniforms this reproduces the reference's sampler tests, which drive the samplers with
 * a generator that always returns its minimum (test/scenario_sampler_test.cpp:198-270). */
int mptg_sample_transform_batch(mptg_ctx* ctx, const mptg_space_desc* space, const double* lo, const double* hi,
                                const void* uniforms, uint32_t n, void* states_out);

/* --------------------------------------------------------- device-resident PRRT
 * The tree (node states, parent indices), its nearest-neighbour structure and every stage of
 * Worker::addSample (src/mpt/impl/prrt/prrt.hpp:411-452) stay on the GPU.  One wave = n_samples iterations of
 * the reference's loop run as a batch: sample (goal-biased until the first goal node, :365-387) -> nearest (:416)
 * -> drop d == 0 (:427) -> steer to `range` (:430-434) -> scenario.valid (:439) -> scenario.link (:441) ->
 * append the survivors in sample order with parent = nearest node (:444-447) -> goal test (:442, GoalState
 * semantics of src/mpt/goal_state.hpp:64-69).  Samples of one wave do not see each other's nodes. */
typedef struct mptg_prrt mptg_prrt;
typedef struct mptg_prrt_params {
    const mptg_space_desc* space;
    const double* lo;       /* sampling bounds, one double per scalar (see mptg_sample_batch) */
    const double* hi;
    double range;           /* Planner::setRange(); use HUGE_VAL for none */
    double goal_bias;       /* Planner::setGoalBias(), default of the reference 0.01 */
    const void* goal_state; /* GoalState: state (space scalar type) or NULL for no goal */
    double goal_radius;
    double link_step;       /* DiscreteMotionValidator step size (mesh geometries) */
    uint64_t seed;
    uint32_t capacity;      /* maximum number of nodes */
    uint32_t max_wave;      /* maximum samples per wave */
} mptg_prrt_params;
int mptg_prrt_create(mptg_ctx* ctx, mptg_geom* geom, const mptg_prrt_params* params, mptg_prrt** out);
int mptg_prrt_destroy(mptg_prrt* prrt);
/* Planner::addStart(state) (prrt.hpp:176-190) */
int mptg_prrt_add_start(mptg_prrt* prrt, const void* state);
/* One wave.  size_out: nodes in the tree afterwards (Planner::size()); goal_node_out: index of the first node
 * that reached the goal, MPTG_NO_INDEX while unsolved (Planner::solved()). */
int mptg_prrt_wave(mptg_prrt* prrt, uint32_t n_samples, uint32_t* size_out, uint32_t* goal_node_out);
uint32_t mptg_prrt_size(const mptg_prrt* prrt);
uint64_t mptg_prrt_samples_drawn(const mptg_prrt* prrt);
/* Node states (AoS, host) and parent indices (MPTG_NO_INDEX for a start) of nodes first .. first+count-1:
 * what Planner::solution() and visitGraph() walk (prrt.hpp:232-262). */
int mptg_prrt_get_tree(mptg_prrt* prrt, uint32_t first, uint32_t count, void* states_out, uint32_t* parents_out);

/* --------------------------------------------------------- device-resident PPRM
 * The roadmap (node states, per-node edge rows, union-find components, start / goal marks), its nearest-neighbour
 * structure and every stage of PPRM's Worker::addSample (src/mpt/impl/pprm/pprm.hpp:298-339) stay on the GPU.
 * One wave = n_samples iterations of the reference's loop (:368-378) run as a batch: sample -> scenario.valid (:299)
 * -> k nearest, k = ceil(kRRG * ln(size + 1)), kRRG = e + e / dimensions (:146,:302-304) -> drop samples closer than
 * epsilon to a node (:306-308) -> scenario.link(sample, neighbour) for every neighbour (:325-326) -> append the nodes
 * in sample order, record the valid edges, merge components (:327-334,:341-362), goal test (:312-316) -> solved when a
 * component holds a start and a goal (impl/pprm/component.hpp).  Samples of one wave do not see each other's nodes.
 * Edge rows: node i keeps the edges found when it was added (to older nodes), `row_stride` slots, unused slots
 * MPTG_NO_INDEX; the roadmap is the undirected union of the rows. */
typedef struct mptg_pprm mptg_pprm;
typedef struct mptg_pprm_params {
    const mptg_space_desc* space;
    const double* lo;       /* sampling bounds, one double per scalar (see mptg_sample_batch) */
    const double* hi;
    const void* goal_state; /* GoalState: state (space scalar type) or NULL: only mptg_pprm_add_state marks goals */
    double goal_radius;
    double link_step;       /* DiscreteMotionValidator step size (mesh geometries) */
    uint64_t seed;
    uint32_t capacity;      /* maximum number of nodes */
    uint32_t max_wave;      /* maximum samples per wave */
    uint32_t max_k;         /* cap on k (0: min(MPTG_MAX_K, k at `capacity` nodes)) = row stride */
} mptg_pprm_params;
#define MPTG_PPRM_START 1u
#define MPTG_PPRM_GOAL 2u
int mptg_pprm_create(mptg_ctx* ctx, mptg_geom* geom, const mptg_pprm_params* params, mptg_pprm** out);
int mptg_pprm_destroy(mptg_pprm* pprm);
/* PPRM-IRS (src/mpt/pprm_irs.hpp:49-99; impl/pprm_irs/pprm_irs.hpp:350-368, shortest_path_check.hpp:111-225): from now on a
 * validated edge (sample, neighbour) of length d is recorded -- as a SPARSE edge -- only if the sparse roadmap does not
 * already join its ends by a path shorter than stretch_weight * d (Planner::setStretchWeight, default 5 in the reference);
 * the edge rows, components, solved() and mptg_pprm_get_graph then describe the sparse roadmap.  Each new node runs the
 * reference's bounded Dijkstra search against the roadmap as it stood when the wave began plus its own kept edges
 * (neighbours nearest first).  Call before the first state is added.  search_capacity: nodes one search may label at
 * first (0: what fits 1 GiB of working storage for a full wave, 256 .. 4096); a wave in which a search needs more is run
 * again with four times the storage (up to one entry per node of `capacity` and 32 GiB in all -- in spaces of many
 * dimensions a search within stretch_weight * d reaches most of the roadmap, for the reference as for this library),
 * beyond that the wave fails with MPTG_ERR_CAPACITY.  keep_dense_edges<true> is a host-planner option (planner.hpp). */
int mptg_pprm_set_spanner(mptg_pprm* pprm, double stretch_weight, uint32_t search_capacity);
/* Planner::addStart(state) / addGoal(state) (pprm.hpp:156-169): addSample with the start / goal mark.
 * node_out: the new node, MPTG_NO_INDEX when the state is invalid or closer than epsilon to a node. */
int mptg_pprm_add_state(mptg_pprm* pprm, const void* state, uint32_t marks, uint32_t* node_out);
/* One wave.  size_out: nodes afterwards (Planner::size()); solved_out: 1 once a start and a goal are connected. */
int mptg_pprm_wave(mptg_pprm* pprm, uint32_t n_samples, uint32_t* size_out, uint32_t* solved_out);
uint32_t mptg_pprm_size(const mptg_pprm* pprm);
uint64_t mptg_pprm_samples_drawn(const mptg_pprm* pprm);
uint32_t mptg_pprm_row_stride(const mptg_pprm* pprm);
/* Nodes first .. first+count-1: states (AoS), edge rows (neighbour index / distance, count * row_stride each), marks
 * (MPTG_PPRM_START | MPTG_PPRM_GOAL) and component representative: what Planner::solution() (Dijkstra over the
 * roadmap, pprm.hpp:218-246) and visitGraph() (:380-387) walk.  Any output may be NULL. */
int mptg_pprm_get_graph(mptg_pprm* pprm, uint32_t first, uint32_t count, void* states_out, uint32_t* edge_idx_out, void* edge_dist_out,
                        uint8_t* marks_out, uint32_t* component_out);

/* --------------------------------------------------------- device-resident PRRT* 
 * Tree (states, parents, costs), nearest-neighbour structure and every stage of PRRT*'s Worker::addSample
 * (src/mpt/impl/prrt_star/prrt_star.hpp:510-657) on the GPU, one wave = n_samples iterations run as a batch against the
 * tree as it stood when the wave began: sample -> nearest -> steer (distance recomputed, :526-537) -> valid -> link ->
 * k nearest (k = ceil(kRRG ln(size+1)), kRRG = rewire_factor e (1 + 1/d), rrg_rewire_neighbors.hpp:53-67) -> parent =
 * first valid neighbour in cost + distance order, up to the near node or the cost cut-off (:565-605) -> append ->
 * rewire every unchecked neighbour whose cost would drop (:626-656).  Inside a wave rewiring offers are evaluated on
 * the costs before the wave's rewiring step, each node takes its best valid offer (smallest cost, then smallest
 * (sample, slot)), all are applied at once (this cannot close a cycle) and the decreases are pushed to the subtrees
 * (:664-688).  Parameters as for mptg_prrt_create plus Planner::setRewireFactor() (default of the reference 1.1). */
typedef struct mptg_prrtstar mptg_prrtstar;
int mptg_prrtstar_create(mptg_ctx* ctx, mptg_geom* geom, const mptg_prrt_params* params, double rewire_factor, mptg_prrtstar** out);
/* rewire_r_nearest (rrg_rewire_neighbors.hpp:102-128): the neighbourhood of a new node is everything within
 * r(n) = r_rrg (ln(n+1) / (n+1))^(1/d), at most MPTG_MAX_K nodes, instead of the k nearest.  r_rrg = rewire_factor
 * (2 (1 + 1/d) measure / unit_ball_d)^(1/d) is computed by the caller, who knows the measure of the sampled region.
 * Before the first wave only. */
int mptg_prrtstar_set_rewire_radius(mptg_prrtstar* star, double r_rrg);
int mptg_prrtstar_destroy(mptg_prrtstar* star);
int mptg_prrtstar_add_start(mptg_prrtstar* star, const void* state);
/* goal_node_out: the goal node of smallest cost so far (Planner::solution() / solutionCost(), :317-337), MPTG_NO_INDEX
 * while unsolved.  Waves of up to 1,024 samples are queued on the device without host synchronisation between their steps
 * (one at the end); the tree is the same either way.  Environment: MPTG_STAR_QUEUED_MAX=n moves that limit (0: never),
 * MPTG_STAR_TIMING=1 prints the host time spent issuing / waiting per queued wave when the planner is destroyed. */
int mptg_prrtstar_wave(mptg_prrtstar* star, uint32_t n_samples, uint32_t* size_out, uint32_t* goal_node_out);
uint32_t mptg_prrtstar_size(const mptg_prrtstar* star);
uint64_t mptg_prrtstar_samples_drawn(const mptg_prrtstar* star);
uint64_t mptg_prrtstar_rewires(const mptg_prrtstar* star);
/* states, parents (MPTG_NO_INDEX for a start) and path costs (space scalar type) of nodes first .. first+count-1 */
int mptg_prrtstar_get_tree(mptg_prrtstar* star, uint32_t first, uint32_t count, void* states_out, uint32_t* parents_out, void* costs_out);

#ifdef __cplusplus
}
#endif
#endif /* MPTG_H */
