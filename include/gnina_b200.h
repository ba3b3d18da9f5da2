/*
 * gnina_b200.h — C ABI of the B200-native CNN-scoring hot path (libgnina_b200.so).
 *
 * Plain pointers and sizes only; no exceptions, torch or C++ types cross this boundary.  Every entry point
 * returns 0 on success and a non-zero gb_status otherwise; gb_last_error() gives the message for the calling
 * thread.  Each declaration cites the reference interface it replaces (paths relative to the gnina tree).
 *
 * Threading model (mirrors the reference): a gb_cnn handle is used by ONE host thread at a time (the reference
 * serialises CNNTorchScorer::score with a recursive_mutex, lib/cnn_torch_scorer.cpp:106, and gives every worker
 * thread its own fresh_copy(), main/main.cpp:1438).  gb_cnn_clone() is the fresh_copy() equivalent; clones share
 * the read-only gb_model weights on the device.  All device work of a handle is ordered on its own CUDA stream.
 */
#ifndef GNINA_B200_H_
#define GNINA_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  GB_OK = 0,
  GB_ERR_USAGE = 1,    /* reference: usage_error (bad model name/file), lib/cnn_torch_scorer.cpp:71,87 */
  GB_ERR_INTERNAL = 2, /* reference: internal_error / VINA_CHECK */
  GB_ERR_CUDA = 3,     /* reference: abort() on CUDA failure, lib/gpu_util.h:20-26 — here reported, never fatal */
  GB_ERR_NO_DEVICE = 4 /* no CUDA device: the product has NO CPU fallback (the reference silently falls back) */
} gb_status;

typedef struct gb_model gb_model; /* one CNN: weights + typing + grid metadata  (TorchModel<isCUDA>, lib/torch_model.h:22-47) */
typedef struct gb_cnn gb_cnn;     /* an ensemble scorer                          (CNNTorchScorer<isCUDA>, lib/cnn_torch_scorer.h:25-57) */

/* GB_ARCH_OVERLAP: the parameter-free "overlay" graph of the reference's own minimisation test
 * (test/gnina/data/overlap.pt, test/gnina/test_min.py): out = (0, mean(rec * lig)), loss = -log out[1]. */
enum { GB_ARCH_DEFAULT2018 = 1, GB_ARCH_DENSE = 2, GB_ARCH_DEFAULT2017 = 3, GB_ARCH_OVERLAP = 4 };
enum { GB_PRECISION_FP32 = 0, /* validation mode: fp32 CUDA-core kernels, reference-layout fp32 grid      */
       GB_PRECISION_FP16_TC = 1 /* fast mode: fused pooled fp16 grid + tcgen05 fp16 convs, fp32 accumulate */ };

typedef struct {
  int32_t arch;            /* GB_ARCH_*                                                      */
  int32_t n_rec_channels;  /* lines of recmap   (lib/torch_model.cpp:110-113)                */
  int32_t n_lig_channels;  /* lines of ligmap                                                */
  int32_t grid_points;     /* per axis: round(dimension/resolution)+1 (48)                   */
  float resolution, dimension, radius_scaling; /* JSON metadata, lib/torch_model.cpp:57-106  */
  int32_t apply_logistic_loss, skip_softmax;
  char name[64];
} gb_model_info;

const char* gb_last_error(void);
const char* gb_version(void);

/* initializeCUDA(device), lib/dl_scorer.cpp:4-33: select the device for the calling thread; returns a
 * cudaError_t-style code (0 ok).  Number of visible devices via gb_device_count. */
int gb_initialize_cuda(int device);
int gb_device_count(void);

/* TorchModel ctor (lib/torch_model.cpp:49-118): read one model.  `path` is a GNB200W1 blob produced from the
 * reference's embedded .pt by tools/extract_models.py (weights + JSON metadata + rec/lig type maps).
 * Unknown file / malformed blob -> GB_ERR_USAGE ("Could not read torch model <name>"). */
int gb_model_load(const char* path, int device, gb_model** out);
int gb_model_load_mem(const void* data, size_t nbytes, int device, gb_model** out);
int gb_model_get_info(const gb_model* m, gb_model_info* info);
void gb_model_release(gb_model* m); /* reference-counted; handles keep their models alive */

/* make_coordset / FileMappedGninaTyper::get_int_type (lib/torch_model.cpp:120-142): smina type -> (channel or
 * -1, xs_radius of the original type).  is_ligand selects ligmap and offsets channels by n_rec_channels
 * (lib/torch_model.cpp:168).  Host-only; used by parity tests. */
int gb_model_type_atoms(const gb_model* m, int is_ligand, const int32_t* smina_type, int n, int32_t* channel,
                        float* radius);

/* CNNTorchScorer ctor (lib/cnn_torch_scorer.cpp:24-92) after name expansion: build an ensemble scorer over
 * n_models models on `device`. */
int gb_cnn_create(gb_model* const* models, int n_models, int device, gb_cnn** out);
/* fresh_copy() (lib/cnn_torch_scorer.h:54): independent handle for another thread; shares device weights and
 * the current receptor. */
int gb_cnn_clone(const gb_cnn* h, gb_cnn** out);
void gb_cnn_destroy(gb_cnn* h);
int gb_cnn_num_models(const gb_cnn* h);

/* Options: "precision" (GB_PRECISION_*), "max_batch" (poses per device pass), "profile" (0/1),
 * "cnn_rotation" (0..24, gnina's --cnn_rotation: every model is evaluated on the unrotated pose and on
 * cnn_rotation - 1 random rotations of receptor + ligand about the grid centre, cnn_torch_scorer.cpp:127-163 ->
 * TorchModel::forward(rotate) lib/torch_model.cpp:170-173; all model x rotation evaluations are ensemble members),
 * "rotation_seed" (gnina's --seed).  The random stream is this library's (libmolgrid's is not reproducible here):
 * gb_cnn_get_rotation returns the matrix used, so callers can reproduce any evaluation exactly. */
int gb_cnn_set_option(gb_cnn* h, const char* key, double value);
double gb_cnn_get_option(const gb_cnn* h, const char* key);

/* DLScorer::setReceptor (lib/dl_scorer.cpp:93-193): upload receptor atoms ONCE (the reference re-types and
 * re-uploads them for every pose, lib/torch_model.cpp:159,181).  xyz is n x 3 floats, smina_type the 28-value
 * smina enum (lib/atom_constants.h:45-75). */
int gb_cnn_set_receptor(gb_cnn* h, const float* xyz, const int32_t* smina_type, int n);

/* The batch form of CNNTorchScorer::score(model&, false, aff, loss, var) (lib/cnn_torch_scorer.cpp:105-198) +
 * TorchModel::forward (lib/torch_model.cpp:153-224) for n_poses ligand poses against the current receptor.
 *   lig_xyz / lig_type : concatenated atoms of all poses (movable ligand atoms INCLUDING polar hydrogens, as
 *                        DLScorer::setLigand passes them, lib/dl_scorer.cpp:72-87)
 *   pose_offsets       : n_poses+1 prefix offsets into those arrays (ragged poses allowed)
 *   centers            : n_poses x 3 grid centres, or NULL => mean of the pose's ligand atoms
 *                        (lib/torch_model.cpp:163-166; --cnn_center otherwise)
 * Outputs (host arrays of n_poses floats, any may be NULL): ensemble means of CNNscore, CNNaffinity, loss and
 * the population variance of the affinities (0 for one model), exactly the four values score() yields. */
int gb_cnn_score_batch(gb_cnn* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets,
                       int n_poses, const float* centers, float* score, float* affinity, float* loss,
                       float* variance);

/* Same, but un-averaged: outputs are [n_models][n_poses] (TorchModel::forward's {pose, affinity, loss}). */
/* with cnn_rotation R > 1 the rows are the model x rotation evaluations: row = model * R + rotation */
int gb_cnn_score_batch_models(gb_cnn* h, const float* lig_xyz, const int32_t* lig_type,
                              const int32_t* pose_offsets, int n_poses, const float* centers, float* pose,
                              float* affinity, float* loss);

/* Row-major 3x3 rotation applied about the grid centre for rotation index `rotation` (0 = identity) of the staged
 * pose `pose` (index within the batch). */
int gb_cnn_get_rotation(const gb_cnn* h, int rotation, int pose, float* matrix9);

/* CNNTorchScorer::score(model&, compute_gradient = true, ...) in batch form: the four outputs as above plus the
 * gradient of the (ensemble-mean) loss with respect to every ligand atom passed, dlig_xyz[n_atoms][3]
 * (TorchModel::forward with autograd + GridMaker::backward, lib/torch_model.cpp:197-221; accumulation and 1/cnt
 * scaling lib/cnn_torch_scorer.cpp:164-179).  The reference adds this to m.minus_forces (lib/model.cu:247-259);
 * untyped atoms (hydrogens) get 0.  drec_xyz (NULL, or [n receptor atoms][3] in gb_cnn_set_receptor's order):
 * getReceptorGradient (lib/torch_model.cpp:226-232), what the reference adds to the flexible residues' atoms; the
 * reference scores one pose per call, and with a shared receptor the quantity only exists per pose: n_poses must be 1.
 * default2018 and dense families.  Option "precision" selects the kernels: GB_PRECISION_FP16_TC (default; ensembles
 * of default2018-architecture models only -- an ensemble with a dense member uses the fp32 kernels) runs forward AND
 * backward on the tensor cores (fp16 gradients with loss scaling, gb_cnn_tc_grad.cu; max |error| 5e-3 of the largest
 * gradient component on the reference's vectors), GB_PRECISION_FP32 the fp32 validation kernels (4e-7). */
int gb_cnn_score_grad(gb_cnn* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets,
                      int n_poses, const float* centers, float* score, float* affinity, float* loss, float* variance,
                      float* dlig_xyz, float* drec_xyz);

/* Split form used for device-resident measurement: stage = host->device copy of the poses (pinned staging,
 * async on the handle's stream); run = all kernels for the staged poses (no host<->device traffic);
 * fetch = device->host copy of the per-pose ensemble results and stream sync. */
int gb_cnn_stage_poses(gb_cnn* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets,
                       int n_poses, const float* centers);
int gb_cnn_run_staged(gb_cnn* h);
int gb_cnn_fetch(gb_cnn* h, float* score, float* affinity, float* loss, float* variance);
/* the same four arrays [4][n_staged] (score, affinity, loss, variance) copied device-to-device into a caller-owned DEVICE
 * buffer on the handle's stream (then synchronised): the multi-GPU score gather (NCCL) reads the results where they are,
 * without a host bounce */
int gb_cnn_fetch_device(gb_cnn* h, float* device_dst);
void* gb_cnn_stream(gb_cnn* h);            /* cudaStream_t of the handle (for CUDA-event timing)            */
int64_t gb_cnn_kernel_launches(gb_cnn* h); /* kernels launched by this handle so far                        */

/* Per-kernel CUDA-event timing (no reference counterpart; measurement only).  Enable with option "profile"=1;
 * read entry `index` (0,1,...) until a non-zero return: kernel-class name, accumulated ms, launches. */
int gb_cnn_profile_read(gb_cnn* h, int index, char* name, int name_cap, double* total_ms, int64_t* count);
int gb_cnn_profile_reset(gb_cnn* h);

/* Test-only (no reference counterpart): raw copy of an internal fast-path buffer of the most recent pass on the
 * calling thread: "x0" (pooled fp16 grid), "x2", "x4" (conv inputs), "y3", "y5" (conv outputs).  out == NULL
 * just reports the size. */
int gb_cnn_debug_read(gb_cnn* h, const char* name, void* out, size_t cap_bytes, size_t* nbytes);

/* GridMaker::forward (libmolgrid; call site lib/torch_model.cpp:181) for parity tests: voxelise the poses for
 * model `model_index` and copy the fp32 grids [n_poses][C][N][N][N] (reference layout, z fastest) to the host. */
int gb_cnn_voxelize(gb_cnn* h, int model_index, const float* lig_xyz, const int32_t* lig_type,
                    const int32_t* pose_offsets, int n_poses, const float* centers, float* grid_out);

/* ------------------------------------------------------------------------------------------------------------
 * smina / Vina empirical scoring (the rows of the hot path that sit next to the CNN in gnina's rescoring loop).
 * Default term set and weights (main/main.cpp:1324-1329): gauss(o=0,w=0.5), gauss(o=3,w=2), repulsion,
 * hydrophobic(0.5,1.5), non_dir_h_bond(-0.7,0), num_tors_div; cutoff 8 A.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct gb_vina gb_vina;

/* precalculate_linear(sf, factor) ctor (lib/precalculate.h:176-208; factor = 32 for docking, main/main.cpp:904-905):
 * tabulates every smina type pair.  weights6 = the 5 term weights + the num_tors_div weight, or NULL for the defaults. */
int gb_vina_create(int device, const float* weights6, float factor, gb_vina** out);
void gb_vina_destroy(gb_vina* h);
/* samples per pair (n = sz(factor * cutoff^2) + 3 = 2051) and one pair's tables: fast[i] (eval_fast, :90-95) and the
 * (e, dor) pairs eval_deriv interpolates (:97-133).  Host-side, for parity tests (V2). */
int gb_vina_table_size(const gb_vina* h);
int gb_vina_prec_table(const gb_vina* h, int t1, int t2, float* fast, float* smooth_e, float* smooth_dor);
/* model::grid_atoms: the rigid receptor atoms; hydrogens are dropped like the reference does. */
int gb_vina_set_receptor(gb_vina* h, const float* xyz, const int32_t* smina_type, int n);
/* cache::populate (lib/cache.cpp:104-184) on the device: one affinity grid of (n[0]+1)x(n[1]+1)x(n[2]+1) points per
 * needed ligand atom type over the box [begin, end] (grid_dims), x fastest like array3d. */
int gb_vina_cache_build(gb_vina* h, const float* begin, const float* end, const int32_t* n, const int32_t* types_needed,
                        int n_types);
int gb_vina_cache_read(gb_vina* h, int type, float* out); /* parity access to one grid */
/* cache::eval / cache::eval_deriv (lib/cache.cpp:50-83 -> grid::evaluate_aux, lib/grid.cpp:96-186) for a batch of
 * poses: trilinear interpolation, curl with cap v, out-of-box penalty slope*miss.  energy[n_poses]; deriv (nullable)
 * [n_atoms][3] = what the reference stores in m.minus_forces. */
int gb_vina_cache_eval(gb_vina* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets, int n_poses,
                       float slope, float v, float* energy, float* deriv);
/* Final scoring of rigid poses: naive_non_cache::eval with precalculate_exact (lib/naive_non_cache.cpp:29-57), i.e.
 * the intermolecular part of model::eval_adjusted (the intramolecular terms cancel, lib/model.cu:401-406), then
 * num_tors_div (lib/everything.h:795-809) -> the "Affinity (kcal/mol)" column.  num_tors per pose as
 * conf_independent_inputs computes it (lib/terms.cpp:74-106); e_inter / affinity: n_poses floats (nullable). */
int gb_vina_score_exact(gb_vina* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets, int n_poses,
                        const float* num_tors, float v, float* e_inter, float* affinity);
/* The DOCKING branch's final score (main/main.cpp:340-344): eval_adjusted with ig = nc_new = non_cache(grid_cache, gd, &prec, slope),
 * i.e. non_cache::eval (lib/non_cache.cpp:52-83) -- coordinates clamped to the search box [box_begin, box_end], pair terms from the
 * SEARCH's precalculate through precalculate::eval (= eval_fast, the piecewise-constant table), curl with vcap, + slope x distance
 * outside the box -- then num_tors_div.  (--score_only and --minimize use gb_vina_score_exact: naive_non_cache + exact terms.) */
int gb_vina_score_noncache(gb_vina* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets, int n_poses,
                           const float* num_tors, float vcap, float slope, const float* box_begin, const float* box_end, float* e_inter,
                           float* affinity);

/* ---- docking inner loop: conformation -> energy/gradient, BFGS, Monte-Carlo chains (one warp per chain) --------
 * The ligand as gnina's model holds it (lib/tree.h, lib/model.h): atoms in the local frame of their torsion-tree
 * segment; segment 0 = rigid root (rigid_body), segments 1.. = torsion segments in DFS pre-order, which is also the
 * order of conf.torsions (heterotree::set_conf, lib/tree.h:361-366); relative origin/axis as segment's ctor stores
 * them (lib/tree.h:208-216); intramolecular interacting pairs (lib/interacting_pairs.h:7-19) and gyration radius
 * (lib/model.cpp:1002-1014).  conf = position[3], orientation quaternion[4] (a,b,c,d), torsions[n_segments-1];
 * change = force[3], torque[3], torsion derivatives. */
typedef struct {
  int32_t n_atoms, n_segments, n_pairs;
  const float* local_xyz;        /* [n_atoms][3] */
  const int32_t* smina_type;     /* [n_atoms]    */
  const int32_t* seg_parent;     /* [n_segments], -1 for the root */
  const int32_t* seg_atom_begin; /* [n_segments] */
  const int32_t* seg_atom_end;
  const float* seg_rel_origin;   /* [n_segments][3] */
  const float* seg_rel_axis;     /* [n_segments][3] */
  const int32_t* pair_a;         /* [n_pairs] */
  const int32_t* pair_b;
  float gyration_radius;         /* model::gyration_radius of the pose the search starts from (heavy atoms about the root origin):
                                  * mutate_conf's rotation amplitude until the first evaluation; after that the chains take it from
                                  * the conformation the model holds, as mutate.cpp:55 does */
} gb_ligand_topology;
/* monte_carlo members (lib/monte_carlo.h:29-41): defaults temperature 1.2, hunt_cap (10, 1.5, 10), min_rmsd 0.5,
 * num_saved_mins 50, mutation_amplitude 2; num_steps and maxiters as main/main.cpp:442-457 derives them. */
typedef struct {
  int32_t num_steps, maxiters, num_saved_mins;
  float temperature, mutation_amplitude, min_rmsd;
  float hunt_cap[3];
} gb_mc_params;

int gb_vina_set_ligand(gb_vina* h, const gb_ligand_topology* lig); /* <= 96 atoms, <= 32 segments */
/* precalculate_splines (lib/precalculate.h:380-449, lib/splines.h:22-138; factor 10 -> 80 intervals per type pair, what
 * --minimize uses, main/main.cpp:1162-1165).  gb_vina_spline_table: the (a,b,c,d) coefficients of one pair
 * [gb_vina_spline_size()][4]; gb_vina_set_precalc(h, 1) makes the intramolecular pair terms of eval_deriv / bfgs / mc use
 * the splines instead of the linear tables. */
int gb_vina_spline_size(const gb_vina* h);
int gb_vina_spline_table(const gb_vina* h, int t1, int t2, float* abcd);
int gb_vina_set_precalc(gb_vina* h, int use_splines);
/* model::eval_deriv with ig = cache (lib/model.cu:202-225): e[n], change[n][6+T], coords[n][n_atoms][3] (nullable).
 * v3 = the curl caps (ligand pairs, grid, other pairs), slope = out-of-box penalty slope of the cache. */
int gb_vina_eval_deriv(gb_vina* h, const float* confs, int n, const float* v3, float slope, float* e, float* change,
                       float* coords);
/* quasi_newton::operator() -> bfgs with fast_line_search (lib/quasi_newton.cpp:49-83, lib/bfgs.h:358-502), n
 * independent minimisations; confs are updated in place. */
int gb_vina_bfgs(gb_vina* h, float* confs, int n, int maxiters, const float* v3, float slope, float* e, float* change,
                 int32_t* n_evals);
/* parallel_mc::operator() / monte_carlo::operator() (lib/parallel_mc.cpp:183-214, lib/monte_carlo.cpp:99-148): n_chains
 * independent Monte-Carlo chains (seeds as parallel_mc draws them, :197-199), each returning up to num_saved_mins
 * RMSD-distinct minima sorted by energy: out_e[n_chains][S], out_conf[n_chains][S][7+T], n_out[n_chains]. */
int gb_vina_mc(gb_vina* h, const gb_mc_params* params, const float* corner1, const float* corner2, const uint32_t* seeds,
               int n_chains, float slope, float* out_e, float* out_conf, int32_t* n_out);
/* Same, additionally recording every chain's current energy after each Monte-Carlo step (monte_carlo.cpp:99-148's
 * `tmp.e`): trace[n_chains][num_steps], nullable.  Test/diagnostic access: chains driven by the same generator must
 * accept the same moves as the CPU restatement. */
int gb_vina_mc_traced(gb_vina* h, const gb_mc_params* params, const float* corner1, const float* corner2, const uint32_t* seeds,
                      int n_chains, float slope, float* out_e, float* out_conf, int32_t* n_out, float* trace);
/* model::eval_deriv with ig = non_cache (lib/non_cache.cpp:126-174): direct sums over the receptor's heavy atoms (the
 * reference pre-selects them with an szv_grid, same order) instead of the affinity grids; [box_begin, box_end] is
 * the grid_dims box whose faces clamp the atom and charge slope x distance outside (check_bounds_deriv, :102-123). */
int gb_vina_eval_deriv_noncache(gb_vina* h, const float* confs, int n, const float* v3, float slope, const float* box_begin,
                                const float* box_end, float* e, float* change);
/* The empirical term non_cache_cnn::eval_deriv mixes into the CNN forces (lib/non_cache_cnn.cpp:113-140, options
 * mix_emp_force / mix_emp_energy / empirical_weight): for every heavy atom the precalculate::eval_deriv sum over the
 * receptor atoms within the cut-off, taken at the atom's position clamped to [box_begin, box_end], then curl(e, deriv, v).
 * e[n_atoms], deriv[n_atoms][3]; hydrogens give zeros; the caller adds its out-of-box penalties. */
int gb_vina_noncache_atoms(gb_vina* h, const float* xyz, const int32_t* smina_type, int n_atoms, const float* box_begin,
                           const float* box_end, float v, float* e, float* deriv);
/* refine_structure (main/main.cpp:131-171) for n conformations at once: up to five quasi-Newton runs
 * (lib/quasi_newton.cpp:49-83, bfgs with fast_line_search) against non_cache with the out-of-box slope 10, 100, ... until
 * every heavy atom is inside the box (non_cache::within, margin 1e-4).  confs are refined in place; e[n] = the last
 * run's energy, or max float when the pose never entered the box (the reference sets out.e = max_fl); within[n],
 * n_evals[n] nullable.  Needs gb_vina_set_receptor + gb_vina_set_ligand, no cache. */
int gb_vina_refine(gb_vina* h, float* confs, int n, int maxiters, const float* v3, const float* box_begin, const float* box_end,
                   float* e, int32_t* within, int32_t* n_evals);
/* minimization_params (lib/common.h:50-60) for the --minimize flows: BFGSAccurateLineSearch (bfgs.h:107-180; main/main.cpp:1160,1186
 * select it, with maxiters 10000 when the user gives none, :1157-1158) and --minimize_early_term (bfgs.h:455-462).  With both flags 0
 * these two calls equal gb_vina_bfgs / gb_vina_refine (fast_line_search, what the docking search uses). */
typedef struct { int32_t maxiters, accurate_line_search, early_term; } gb_minimization_params;
/* quasi_newton::operator() (lib/quasi_newton.cpp:49-83) with these parameters on the cache field: confs in/out */
int gb_vina_minimize(gb_vina* h, float* confs, int n, const gb_minimization_params* mp, const float* v3, float slope, float* e,
                     float* change, int32_t* n_evals);
/* refine_structure (main/main.cpp:131-171) with these parameters: the --local_only / --minimize branch (:264-268) */
int gb_vina_refine_minimize(gb_vina* h, float* confs, int n, const gb_minimization_params* mp, const float* v3, const float* box_begin,
                            const float* box_end, float* e, int32_t* within, int32_t* n_evals);
/* merge_output_containers over the chains' containers (lib/parallel_mc.cpp:165-181 -> add_to_output_container,
 * lib/coords.cpp:43-56, find_closest / rmsd_upper_bound :24-41): chains in order, each chain's minima in order;
 * a pose closer than min_rmsd (RMSD over the n_atoms coordinates given) to a kept one replaces it if better,
 * otherwise it is appended while there is room or replaces the worst; the container is re-sorted by energy after
 * every insertion.  Host-only (no device work).  e[n_chains][S], coords[n_chains][S][n_atoms][3], n_out[n_chains];
 * kept[max_size] receives flat indices chain * S + k in final (energy) order, *n_kept their number. */
int gb_vina_merge_outputs(const float* e, const float* coords, const int32_t* n_out, int n_chains, int S, int n_atoms,
                          float min_rmsd, int max_size, int32_t* kept, int32_t* n_kept);

#ifdef __cplusplus
}
#endif
#endif /* GNINA_B200_H_ */
