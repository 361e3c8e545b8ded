/*
 * peprml.h -- C ABI of the B200-native maximum-likelihood engine that replaces the raxmlHPC / raxmlHPC-PTHREADS /
 * FastTree_WAG executions PEPR performs through its tool runners.
 *
 * The reference crosses into native code ONLY by fork/exec with files in the CWD
 * (src/edu/vt/vbi/ci/pepr/util/ExecUtilities.java:23-42,80-119); there is no FFI to copy.  Each entry point below
 * names the reference interface whose work it takes over.  All functions return 0 on success and a negative
 * PML_E* code on failure; the message is available from pml_last_error() of the context that was used.
 * Plain pointers and sizes only; no C++ / torch types.  All arrays are caller-allocated.
 *
 * Threading: a pml_ctx owns one GPU, one CUDA stream and (optionally) one NCCL rank.  Distinct contexts may be
 * used from distinct host threads concurrently (PEPR runs up to `tree_threads` runners in one JVM,
 * PhylogenomicPipeline2.java:1233-1254); one context must not be used from two threads at once.
 * There is no CPU fallback: without a usable CUDA device pml_ctx_create fails with PML_ENODEVICE.
 */
#ifndef PEPRML_H
#define PEPRML_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PML_OK 0
#define PML_EINVAL (-1)    /* bad argument / malformed newick / unknown taxon */
#define PML_ENODEVICE (-2) /* no CUDA device, or CUDA runtime error (see pml_last_error) */
#define PML_ENOMEM (-3)
#define PML_ECOMM (-4)     /* NCCL error, or a rank of the group did not deliver its sums in time */
#define PML_ESTATE (-5)    /* call order (e.g. evaluate before model_set) */

#define PML_UNIQUE_ID_BYTES 128

typedef struct pml_ctx pml_ctx;
typedef struct pml_aln pml_aln;
typedef struct pml_tree pml_tree;

const char *pml_version(void);

/* ---- context: one GPU (+ one rank of a site-sharded group) -------------------------------------------------
 * Replaces `-T n` of raxmlHPC-PTHREADS (RAxMLRunner.java:130-132): the pthread workers over alignment patterns
 * become `nranks` contexts over pattern shards; the masterBarrier reduction becomes an NCCL allreduce of 1-3 doubles.
 * nranks == 1: unique_id may be NULL.  nranks > 1: every rank passes the same id obtained from pml_comm_unique_id
 * on one rank (distributed by the host: a Java array shared by threads, torch.distributed broadcast, a file ...). */
int pml_comm_unique_id(unsigned char id[PML_UNIQUE_ID_BYTES]);
int pml_ctx_create(int gpu_id, int rank, int nranks, const unsigned char *unique_id, pml_ctx **out);
/* All ranks of a site-sharded group inside ONE process: what `-T n` is to raxmlHPC-PTHREADS when PEPR's single JVM starts it
 * (RAxMLRunner.java:130-132).  out receives ngpu contexts, rank i on gpu_ids[i] (distinct GPUs).  The in-kernel reduction
 * reaches the peers' mailboxes through cudaDeviceEnablePeerAccess (no CUDA IPC, which cannot map a handle inside the process
 * that exported it), so pml_ctx_collective() is 2 here exactly as it is for one process per GPU.  Each context is driven
 * by its own host thread and all threads make the same calls in the same order (a branch pass waits for its peers' sums --
 * for at most PEPRML_PEER_TIMEOUT_MS, default 10 s, then the call fails with PML_ECOMM; nothing hangs).  pml_aln_load on
 * such contexts sorts the alignment once for the whole group.  Destroy every context with pml_ctx_destroy when all are idle. */
int pml_group_create(const int *gpu_ids, int ngpu, pml_ctx **out);
void pml_ctx_destroy(pml_ctx *);
const char *pml_last_error(const pml_ctx *); /* ctx may be NULL: last creation error of this thread */
int pml_ctx_sync(pml_ctx *);
/* how the 1-3 doubles of a branch pass are summed over the ranks: 0 = single rank, 1 = NCCL allreduce on the stream,
 * 2 = inside the branch kernel over NVLink peer memory (CUDA IPC mailboxes; the default whenever peers can be mapped) */
int pml_ctx_collective(const pml_ctx *);

/* ---- alignment -------------------------------------------------------------------------------------------
 * Replaces the phylip hand-off SequenceAlignment.getAlignmentAsExtendedPhylipUsingTaxonNames -> `-s file`
 * (SequenceAlignment.java:489-522; RAxMLRunner.java:100-107) and raxmlHPC's getinput/makevalues/sitesort:
 * chars = ntax x nsites raw residue letters, row-major; letters ARNDCQEGHILKMFPSTWYV (any case), B, Z; every other
 * byte (X ? * - ...) is "undetermined".  site_weights (NULL = all 1) replaces `-a weightFile`.
 * Identical columns are merged (lexicographic column sort in taxon order) exactly as the reference does, so pattern
 * order and pattern weights equal raxmlHPC's.  With nranks > 1 each rank passes the FULL alignment and keeps its
 * contiguous block of patterns on its GPU. */
int pml_aln_load(pml_ctx *, int ntax, int64_t nsites, const char *const *names, const uint8_t *chars,
                 const int32_t *site_weights, pml_aln **out);
int pml_aln_load_phylip(pml_ctx *, const char *path, const char *weights_path /* NULL */, pml_aln **out);
void pml_aln_free(pml_aln *);
int pml_aln_dims(const pml_aln *, int *ntax, int64_t *nsites, int64_t *npatterns, int64_t *npatterns_local);
/* weights: npatterns int32 (global pattern order); site_to_pattern: nsites int64 (-1 for columns of weight 0) */
int pml_aln_patterns(const pml_aln *, int32_t *weights, int64_t *site_to_pattern);
const char *pml_aln_name(const pml_aln *, int taxon);
/* host-only pattern crunch (no GPU, no context): the same code pml_aln_load runs.  codes_out: ntax x nsites bytes
 * capacity, filled as ntax x *npatterns row-major; weights_out / site_to_pattern as above. */
int pml_crunch_patterns(int ntax, int64_t nsites, const uint8_t *chars, const int32_t *site_weights, uint8_t *codes_out,
                        int32_t *weights_out, int64_t *site_to_pattern, int64_t *npatterns);

/* host-only: the same crunch as `nranks` ranks of a multi-process group perform it -- the radix sort split by first-residue
 * bucket, the sorted column order exchanged by an element-wise sum (an NCCL allreduce inside pml_aln_load; threads and a
 * barrier here) -- with every rank's code block put side by side.  Must equal pml_crunch_patterns bit for bit. */
int pml_crunch_patterns_sharded(int nranks, int ntax, int64_t nsites, const uint8_t *chars, const int32_t *site_weights,
                                uint8_t *codes_out, int32_t *weights_out, int64_t *site_to_pattern, int64_t *npatterns);

/* ---- model: `-m PROTGAMMAWAG` (PhylogenomicPipeline2.java:248-250) -----------------------------------------
 * WAG exchangeabilities + fixed WAG frequencies, 4 mean-Gamma categories with shape alpha. */
int pml_model_set(pml_aln *, const char *model, double alpha);
int pml_model_get(const pml_aln *, double *alpha, double rates4[4]);
/* host-only helpers (no GPU needed) used by data generators and tests */
int pml_wag_pmatrix(double t, double rate, double P[400]);     /* row-major P(i->j) */
int pml_wag_frequencies(double pi[20]);
int pml_gamma_rates(double alpha, int ncat, double *rates);

/* ---- tree ------------------------------------------------------------------------------------------------
 * newick as BasicTree prints / parses it (BasicTree.java:131-409,450-520): rooted-binary or trifurcating top,
 * optional inner labels, optional branch lengths (missing -> 0.1 substitutions/site). Replaces `-t tree`. */
int pml_tree_load(pml_aln *, const char *newick, pml_tree **out);
void pml_tree_free(pml_tree *);
int pml_tree_num_branches(const pml_tree *);
int pml_tree_branch(const pml_tree *, int branch, int *node_a, int *node_b, double *length); /* nodes < ntax are tips */
int pml_tree_set_branch(pml_tree *, int branch, double length);
/* writes the tree as raxmlHPC's RAxML_result file does: trifurcation at the inner node next to the first taxon,
 * lengths with 20 decimals, terminated by ":0.0;" ; returns needed size if cap is too small */
int64_t pml_tree_newick(const pml_tree *, char *buf, size_t cap);

/* ---- likelihood: newview traversal + root evaluate (raxmlHPC newviewGTRGAMMAPROT / evaluateGTRGAMMAPROT) ---
 * Replaces `-f g` / `-f n` scoring (RAxMLRunner.runRaxmlPerSiteLL, RAxMLRunner.java:162-213) at FIXED parameters.
 * weights: NULL = alignment's pattern weights, else npatterns int32 (e.g. one bootstrap replicate).
 * per_site: NULL or nsites doubles in ORIGINAL column order (what PhylogenomicPipeline2.getTreeScore sums, :1482-1500);
 * with nranks > 1 per_site is filled with this rank's patterns only and 0 elsewhere (sum over ranks = full vector). */
int pml_evaluate(pml_tree *, const int32_t *weights, double *lnl, double *per_site);
/* forces every inner CLV to be recomputed on the next call (full traversal) */
int pml_tree_invalidate(pml_tree *);
/* counters since tree creation: CLV site-updates by case (0 tip-tip, 1 tip-inner, 2 inner-inner) and kernel launches */
int pml_tree_stats(const pml_tree *, int64_t site_updates[3], int64_t *kernel_launches);
/* Newton-Raphson passes since tree creation that ended in raxmlHPC's bad-curvature retry (z = 0.37 z + 0.63): each one makes
 * the smoothing pipeline drop the branch it had queued speculatively */
int64_t pml_tree_nr_retries(const pml_tree *);

/* ---- topological constraints (FastTreeRunner.java:53-83: `FastTree -constraints file`; encoder :243-273) -----------------------
 * text = FastTree's constraint alignment (">name" + a row of 0 / 1 / - per taxon, one column per split; unlisted taxa are free).
 * While set, pml_tree_start_parsimony grows its tree only through insertions that keep every split possible, and pml_search /
 * pml_bootstrap_trees score only pruning-regrafting moves whose result displays every split.  NULL or "" clears.
 * pml_constraints_from_tree writes the constraint alignment of a tree exactly as getFastTreeConstraintsForTree does (taxa
 * sorted, one column per node); returns the bytes needed including the terminator. */
int pml_aln_set_constraints(pml_aln *, const char *text);
int pml_aln_num_constraints(const pml_aln *);      /* splits that can fail (two or more taxa on either side) */
int pml_tree_satisfies_constraints(const pml_tree *); /* 1 / 0 */
int64_t pml_constraints_from_tree(const char *newick, char *buf, size_t cap);
/* host only (no device): the constrained stepwise-addition parsimony tree, and whether a tree displays every split */
int64_t pml_parsimony_tree_constrained(int ntax, int64_t nsites, const char *const *names, const uint8_t *chars, int64_t seed,
                                       const char *constraints, char *buf, size_t cap, int64_t *score);
int pml_newick_satisfies_constraints(const char *newick, const char *const *names, int ntax, const char *constraints);

/* ---- device-side timing of the engine's own kernels (CUDA events on the context's stream) -------------------
 * Between begin and end every launch of kind k is bracketed by a pair of events; end() synchronises and returns, per
 * kind, the summed device milliseconds, the launch count and the pattern rows processed.
 * A side of an update or of a branch is an inner node (I), a tip (T) or a folded cherry (C: an inner node whose two children
 * on that side are tips; its CLV is formed inside the consuming kernel and never stored).  Kinds:
 *   0..5   CLV update by its two children: TT (stored cherry), TI, II, TC, CC, CI
 *   6..8   branch pass, one end an inner node, the other I / T / C;   9..11  the same as root evaluate (per-pattern lnL kept)
 *   12     Newton-Raphson iteration on a stored product table
 *   13..27 fused CLV update + branch pass: 13 + 3 * (children of the update: TI, II, TC, CC, CI) + (far end: I, T, C)
 * pml_kind_info gives a kind's name, its algorithmic bytes per pattern (CLV rows read and written + tip codes +
 * weights) and the DMMA.8x8x4 it issues per 16-pattern tile and rate-category warp (512 flop each, four warps per tile). */
#define PML_NKINDS 28
int pml_kind_info(int kind, char *name, size_t cap, int *bytes_per_pattern, int *dmma_per_tile);
int pml_profile_begin(pml_ctx *);
int pml_profile_end(pml_ctx *, double ms[PML_NKINDS], int64_t launches[PML_NKINDS], int64_t rows[PML_NKINDS]);

/* profiling aid: while enabled (on = 1: CLV kernel, 4 / 5: CLV kernel with one / no tip child only, 2 / 3: branch kernel with two inner ends / a tip end, 6: fused kernel (inner children, inner far end), 0: off), the kernel accumulates per warp of its
 * first CTA the clock cycles spent in each pipeline phase (12 warps x 8 counters: wait data, fragments + wait turn, MMAs,
 * products, wait slot, store, tiles, prologue) */
int pml_trace_enable(pml_ctx *, int on);
/* profiling aid: between begin and read every CLV / fused launch leaves six %globaltimer stamps (ns): CTA 0 enters, its
 * dependency wait returns, its first MMA turn, its last tile leaves the MMA warps, (fused) the last CTA has drawn its ticket,
 * (fused) the result is published.  read returns the number of launches captured (stamps: n x 6, kinds: n) and ends the capture. */
int pml_timeline_begin(pml_ctx *);
int pml_timeline_read(pml_ctx *, uint64_t *stamps, int32_t *kinds, int cap);
int pml_trace_read(pml_ctx *, int64_t out[96]);

/* host builds of two pieces of the device arithmetic, for CPU tests of what DESIGN.md claims about them (no GPU needed):
 * pml_debug_exp_neg is the exponential the kernels use for lambda_k r_c t <= 0 (pmatrix.cuh: relative error <= 1.4 x 2^-53 on
 * [-708, 0], 0 below); pml_debug_nr_step is raxmlHPC's guarded Newton-Raphson step on a branch of length t with derivatives
 * d1, d2 (topLevelMakenewz), stated on z = exp(-t) (in_t = 0, the reference rule) or as the device runs it, on t (in_t = 1);
 * returns the step's status (1 done, 2 bad curvature: derivatives again at *t_new) and the new length. */
double pml_debug_exp_neg(double x);
int pml_debug_nr_step(double t, double d1, double d2, int in_t, double *t_new);

/* stopwatch on the context's stream: start records an event, stop records another, synchronises and returns the
 * device milliseconds in between (host control flow between the two is included, as it should be for a step time) */
int pml_timer_start(pml_ctx *);
int pml_timer_stop(pml_ctx *, double *ms);

/* ---- Newton-Raphson branch-length derivatives (sumGAMMAPROT + coreGTRGAMMAPROT) ----------------------------
 * lnL and d lnL/dt, d2 lnL/dt2 of `branch` at length t (other parameters fixed). */
int pml_branch_derivs(pml_tree *, int branch, double t, const int32_t *weights, double *lnl, double *d1, double *d2);

/* ---- `-f e`: optimise all branch lengths (+ alpha) on the fixed topology (treeEvaluate/optAlpha/modOpt) -----
 * Replaces RAxMLRunner.runRaxmlParsimonyWithBranchLengths (RAxMLRunner.java:215-280) and
 * FastTreeRunner.getRaxmlBranchLengths (FastTreeRunner.java:142-199).  eps = raxml `-e` (default 0.1 lnL units). */
int pml_optimize(pml_tree *, int opt_alpha, double eps, const int32_t *weights, double *lnl, double *alpha);
/* one smoothing sweep (one guarded NR step on every branch); returns 1 in *converged when no branch moved */
int pml_smooth_branches(pml_tree *, int sweeps, const int32_t *weights, int *converged);

/* ---- tree search: parsimony start tree + lazy subtree-pruning-regrafting hill climbing --------------------------
 * Takes over what `-f d` does inside raxmlHPC (makeParsimonyTree, treeOptimizeRapid/rearrangeBIG/testInsertBIG) for
 * RAxMLRunner.run (RAxMLRunner.java:79-152).  The search is a heuristic: it is compared with the reference by the lnL of
 * the tree it finds, not move by move.
 * pml_tree_start_parsimony: randomised stepwise addition under Fitch parsimony (taxon order from the randum stream) on
 *   the patterns weighted by `weights` (a bootstrap replicate); the per-pattern Fitch scans run on the GPU, one launch per
 *   added taxon, and give exactly the tree of the host-only pml_parsimony_tree.
 * pml_score_spr_candidates: prune inner node `node` together with the subtree behind its neighbour `keep`, insert it
 *   into every branch within `radius` steps (halved branch, unchanged subtree branch: raxmlHPC's lazy insertion) and
 *   return the lnL of each candidate; the tree is left unchanged.  targets/lnl: caller arrays of capacity *ncand in,
 *   number filled out.
 * pml_search: repeat { for every (node, subtree): best lazy candidate -> apply, re-optimise the branches around the
 *   insertion, keep if lnL improves } until a whole round gains < eps or max_rounds is reached. */
int pml_tree_start_parsimony(pml_aln *, int64_t seed, const int32_t *weights /* NULL = alignment's */, pml_tree **out);
/* host-only (no GPU): the same start tree as newick text + its parsimony score */
int64_t pml_parsimony_tree(int ntax, int64_t nsites, const char *const *names, const uint8_t *chars, int64_t seed, char *buf,
                           size_t cap, int64_t *score);
/* applies one pruning-regrafting move for good: `node` and the subtree behind its neighbour `keep` go into branch `target`
 * (one of the branches pml_score_spr_candidates lists); branch ids of unaffected branches are preserved */
int pml_tree_spr(pml_tree *, int node, int keep, int target);
int pml_tree_neighbors(const pml_tree *, int node, int nbr[3]);
int pml_score_spr_candidates(pml_tree *, int node, int keep, int radius, const int32_t *weights, int *targets, double *lnl,
                             int *ncand);
int pml_search(pml_tree *, int radius, int max_rounds, double eps, const int32_t *weights, double *lnl, int *accepted_moves);

/* ---- bootstrap replicates as integer site-weight vectors (computeNextReplicate + randum) -------------------
 * Replaces `-x seed -N reps` / `-b seed` weight generation (RAxMLRunner.java:112-124).  out: nrep x npatterns int32,
 * bit-identical to raxmlHPC for the same seed; *seed is advanced so that consecutive calls continue the stream. */
int pml_bootstrap_weights(const pml_aln *, int64_t *seed, int nrep, int32_t *out);
/* same stream from explicit pattern weights (host-only; no context needed) */
int pml_bootstrap_weights_host(const int32_t *pattern_weights, int64_t npatterns, int64_t *seed, int nrep, int32_t *out);
/* lnL of `nrep` weight vectors on the current tree/parameters in one pass (W: nrep x npatterns, global order) */
int pml_evaluate_replicates(pml_tree *, const int32_t *W, int nrep, double *lnl);

/* ---- replicate trees, sharded by REPLICATE ---------------------------------------------------------
 * Replaces PEPR's support-tree workers (PhylogenomicPipeline2.java:1227-1275: one runner thread per support tree, results
 * collected on the host) and the replicate searches of `-f a -x seed -N nrep` (RAxMLRunner.java:112-124).
 * Runs the replicates first, first + stride, ... (< nrep) on THIS context: weights of replicate r from raxmlHPC's stream
 * (the whole stream is drawn, so r carries the same vector whatever the sharding) -> parsimony start tree on the replicate
 * (seed parsimony_seed + 1 + r) -> alpha + branch lengths (eps 5) -> `rounds` lazy-SPR rounds of `radius` with smoothing.
 * With one single-rank context per GPU holding the FULL pattern set and (first, stride) = (g, ngpu) the replicates need no
 * communication at all; the host concatenates the newick texts.  newicks: the share's trees, one per line, in replicate
 * order (capacity: count x pml_newick_capacity()); lnl / seconds (NULL or nrep entries): filled at the share's indices. */
int64_t pml_newick_capacity(const pml_aln *);
int pml_bootstrap_trees(pml_aln *, int64_t weight_seed, int64_t parsimony_seed, int nrep, int first, int stride, int radius,
                        int rounds, double eps, char *newicks, size_t cap, double *lnl, double *seconds);

/* ---- support estimation (integer, bit exact) --------------------------------------------------------------
 * TreeSupportDecorator.addSupportValues (TreeSupportDecorator.java:86-163) with Bipartition canonical form
 * (Bipartition.java:41-64): writes the main tree with, for every inner node, the NUMBER of support trees that
 * contain the same bipartition.  as_percent != 0 gives raxmlHPC `-f b` labels instead (round-half-up integer %,
 * RAxMLRunner.getSupportDecoratedTree, RAxMLRunner.java:453-516).  Host-only; ctx may be NULL. */
int64_t pml_support_tree(const char *main_newick, const char *const *trees, int ntrees, int as_percent, char *buf,
                         size_t cap);
/* counts per non-trivial split of the main tree in the order the labels appear in pml_support_tree's output */
int pml_support_counts(const char *main_newick, const char *const *trees, int ntrees, int32_t *counts, int *nsplits);

#ifdef __cplusplus
}
#endif
#endif
