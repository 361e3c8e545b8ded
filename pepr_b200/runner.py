"""Host-side mirror of PEPR's tool-runner interface for the maximum-likelihood path, in-process over the C ABI.

It keeps the method names, argument meaning and error behaviour of
  edu.vt.vbi.ci.pepr.tree.RAxMLRunner           (src/edu/vt/vbi/ci/pepr/tree/RAxMLRunner.java)
  edu.vt.vbi.ci.pepr.tree.FastTreeRunner        (.../FastTreeRunner.java:38-135 run, :142-199 getRaxmlBranchLengths)
  edu.vt.vbi.ci.pepr.tree.TreeSupportDecorator  (.../TreeSupportDecorator.java:86-163)
so that the parity tests read like calls PEPR itself makes.  The Java twin (`B200MLRunner`) binds the
same C entry points through JNI / Panama; no JDK exists in this image, so this Python class is the executable mirror.

Error convention of the reference: a failed tool run is logged, the result file is then missing and the getter returns
""/null (ExecUtilities.java:29-35, RAxMLRunner.java:307-315).  Here the failure is kept in `last_error`, logged, and the
getters return "" / None in the same way; `strict=True` raises instead.
"""
import logging

import numpy as np

from . import engine as _e

logger = logging.getLogger("PEPR")

ML_ALGORITHM, PARSIMONY_ALGORITHM, PARSIMONY_WITH_BL_ALGORITHM, PER_SITE_LL_ALGORITHM = range(4)


class SequenceAlignment:
    """the slice of edu.vt.vbi.ci.pepr.alignment.SequenceAlignment the runners use: taxon names + one string per taxon"""

    def __init__(self, names, seqs, name="alignment"):
        if len(names) != len(seqs) or len({len(s) for s in seqs}) != 1:
            raise ValueError("alignment needs one equal-length sequence per taxon")
        self.names, self.seqs, self.name = list(names), list(seqs), name

    def getNTax(self):
        return len(self.names)

    def getLength(self):
        return len(self.seqs[0])

    def getAlignmentAsExtendedPhylipUsingTaxonNames(self):
        """SequenceAlignment.java:489-522"""
        w = max(len(n) for n in self.names) + 1
        return "%d %d\n" % (len(self.names), len(self.seqs[0])) + "".join(n.ljust(w) + s + "\n" for n, s in zip(self.names, self.seqs))


def _topology_only(newick):
    """RAxML_parsimonyTree form: no branch lengths"""
    import re
    return re.sub(r":[0-9.eE+-]+", "", newick)


class B200MLRunner:
    """drop-in for RAxMLRunner on the likelihood path: `-f d`, `-f a`, `-f e`, `-f g`, `-f b`, bootstrap weight vectors.

    `threads` (RAxMLRunner's `-T`) selects nothing here: pattern parallelism is the GPU's.  run() without a start tree
    performs the ML tree search (parsimony start tree + lazy SPR + model optimisation), with bootstrapReps > 0 preceded by
    that many replicate searches whose splits are drawn on the ML tree as integer percentages."""

    def __init__(self, threads=1, gpu=0, ctx=None, strict=False):
        self.threads, self.strict = threads, strict
        self._own_ctx = ctx is None
        self.ctx = ctx if ctx is not None else _e.Context(gpu)
        self.matrix = "PROTGAMMAWAG"            # the pipeline always sets it (PhylogenomicPipeline2.java:248-250)
        self.alignment = None
        self.bootstrap_reps = 0
        self.algorithm = ML_ALGORITHM
        self.per_site_ll_trees = None
        self.start_tree = None
        self.random_seed = 12345
        self.eps = 0.1
        self.last_error = None
        self._best_tree = ""
        self._per_site_lines = None
        self._lnl = None
        self._alpha = None
        self.tree_options = ""                  # what PEPRTracker.setTreeOptions would receive
        self.constraints = None                 # FastTree constraint alignment (B200FastTreeRunner passes it on)
        self.site_weights = None                # integer column weights (raxmlHPC -a): gene-wise jackknife masks use them

    # ---- configuration (RAxMLRunner setters) ------------------------------------------------------------------
    def setAlignment(self, alignment):
        self.alignment = alignment

    def setMatrix(self, m):
        self.matrix = m

    def setBootstrapReps(self, n):
        self.bootstrap_reps = int(n)

    def setThreadCount(self, n):
        self.threads = int(n)

    def setParsimonyWithBL(self, flag=True):
        if flag:
            self.algorithm = PARSIMONY_WITH_BL_ALGORITHM

    def setParsimonyOnly(self, flag=True):
        """`-y`: stop after the parsimony start tree (RAxMLRunner.java:132-138); with bootstrapReps > 0 the reference asks for
        `-Y -N reps` (one parsimony tree per replicate)"""
        if flag:
            self.algorithm = PARSIMONY_ALGORITHM

    def setPerSiteLogLikelihoods(self, flag=True):
        if flag:
            self.algorithm = PER_SITE_LL_ALGORITHM

    def setPerSiteLLTrees(self, trees):
        self.per_site_ll_trees = list(trees)

    def setStartTree(self, newick):
        """the `-t` tree of `-f e` (FastTreeRunner.getRaxmlBranchLengths / the second run of parsimony-with-BL)"""
        self.start_tree = newick

    # ---- execution ------------------------------------------------------------------------------------------------
    def _fail(self, msg):
        self.last_error = msg
        logger.error(msg)
        if self.strict:
            raise _e.EngineError(msg)

    def setSiteWeights(self, weights):
        """integer weight per alignment column (raxmlHPC `-a weightFile`); None = every column once.  A gene-wise jackknife
        replicate is the 0 / 1 mask of the genes it keeps (gene_block_weights) over ONE resident supermatrix instead of a new
        concatenated alignment per replicate"""
        self.site_weights = None if weights is None else np.ascontiguousarray(weights, np.int32)

    def _load(self):
        a = self.alignment
        aln = _e.Alignment(self.ctx, a.names, a.seqs, alpha=1.0, model=self.matrix, site_weights=self.site_weights)
        if self.constraints:
            aln.set_constraints(self.constraints)
        return aln

    def _optimise(self, aln, newick):
        tree = _e.Tree(aln, newick)
        for e in range(tree.num_branches):      # raxmlHPC ignores the lengths of the -t tree and starts at z = 0.9
            tree.set_branch(e, -np.log(0.9))
        aln.set_model(1.0, self.matrix)
        lnl, alpha = tree.optimize(True, self.eps)
        return tree, lnl, alpha

    def run(self):
        self.last_error, self._best_tree, self._per_site_lines = None, "", None
        if self.alignment is None:
            return self._fail("no alignment set")
        self._parsimony_tree, self._parsimony_trees = None, []
        try:
            if self.algorithm == PER_SITE_LL_ALGORITHM:
                self._run_per_site_ll()
            elif self.algorithm == PARSIMONY_WITH_BL_ALGORITHM:
                self._run_parsimony_with_bl()
            elif self.algorithm == PARSIMONY_ALGORITHM:
                self._run_parsimony_only()
            elif self.start_tree is not None:
                self._run_branch_lengths()
            else:
                self._run_search()
        except _e.EngineError as ex:
            self._fail(str(ex))

    def _run_branch_lengths(self):
        """`-f e -m <matrix> -s aln -t tree` (RAxMLRunner.java:245-262, FastTreeRunner.java:145-184)"""
        self.tree_options = "peprml -f e -m %s" % self.matrix
        aln = self._load()
        try:
            tree, self._lnl, self._alpha = self._optimise(aln, self.start_tree)
            self._best_tree = tree.newick()
            tree.close()
        finally:
            aln.close()

    def _run_parsimony_with_bl(self):
        """runRaxmlParsimonyWithBranchLengths (RAxMLRunner.java:215-280): `-f d -y` (parsimony start tree, topology only),
        then `-f e -t RAxML_parsimonyTree.<run>` (alpha + branch lengths on that topology) -> RAxML_result.<run>BL"""
        self.tree_options = "peprml -f e -m %s -t <parsimony tree>" % self.matrix
        aln = self._load()
        try:
            start = _e.Tree(aln, parsimony_seed=self.random_seed)
            self._parsimony_tree = _topology_only(start.newick())
            start.close()
            tree, self._lnl, self._alpha = self._optimise(aln, self._parsimony_tree)
            self._best_tree = tree.newick()
            tree.close()
        finally:
            aln.close()

    def _run_parsimony_only(self):
        """`-f d ... -y`, or `-Y -N reps` when bootstrapReps > 0 (RAxMLRunner.java:132-138): parsimony trees, no likelihood"""
        self.tree_options = "peprml -f d -m %s %s" % (self.matrix, "-Y -N %d" % self.bootstrap_reps if self.bootstrap_reps else "-y")
        aln = self._load()
        try:
            for r in range(max(1, self.bootstrap_reps)):
                t = _e.Tree(aln, parsimony_seed=self.random_seed + r)
                self._parsimony_trees.append(_topology_only(t.newick()))
                t.close()
            self._parsimony_tree = self._parsimony_trees[0]
        finally:
            aln.close()

    def _search_one(self, aln, weights, seed, rounds, final):
        tree = _e.Tree(aln, parsimony_seed=seed, weights=weights)
        aln.set_model(1.0, self.matrix)
        tree.optimize(True, 5.0, weights=weights)
        lnl, _ = tree.search(radius=5, max_rounds=rounds, eps=self.eps, weights=weights)
        alpha = aln.alpha
        if final:
            lnl, alpha = tree.optimize(True, self.eps, weights=weights)
        return tree, lnl, alpha

    def _run_search(self):
        """`-f d` (bootstrapReps == 0) or `-f a -x seed -N reps` (RAxMLRunner.java:112-132)"""
        self.tree_options = "peprml -f %s -m %s" % ("a" if self.bootstrap_reps else "d", self.matrix)
        aln = self._load()
        try:
            boots = []
            if self.bootstrap_reps:
                W, _ = aln.bootstrap_weights(self.random_seed, self.bootstrap_reps)
                for r in range(self.bootstrap_reps):
                    t, _, _ = self._search_one(aln, W[r], self.random_seed + 1 + r, 1, False)
                    boots.append(t.newick())
                    t.close()
            tree, self._lnl, self._alpha = self._search_one(aln, None, self.random_seed, 10, True)
            self._best_tree = tree.newick()
            tree.close()
            self._best_with_supports = _e.support_tree(self._best_tree, boots, as_percent=True) if boots else ""
            self.bootstrap_trees = boots
        finally:
            aln.close()

    def getBestTreeWithSupports(self):
        """RAxML_bipartitions.<run> of `-f a` (RAxMLRunner.java:302-318)"""
        return getattr(self, "_best_with_supports", "")

    def _run_per_site_ll(self):
        """`-f g -z trees` (RAxMLRunner.java:162-213): per tree, optimise then print per-site lnL in column order"""
        self.tree_options = "peprml -f g -m %s" % self.matrix
        trees = self.per_site_ll_trees or []
        aln = self._load()
        lines = ["  %d  %d" % (len(trees), self.alignment.getLength())]
        try:
            for i, nw in enumerate(trees):
                tree, self._lnl, self._alpha = self._optimise(aln, nw)
                _, ps = tree.evaluate(per_site=True)
                lines.append("tr%d\t" % (i + 1) + "".join("%.6f " % v for v in ps))
                self._best_tree = tree.newick()
                tree.close()
        finally:
            aln.close()
        self._per_site_lines = lines

    # ---- results (RAxMLRunner getters) ------------------------------------------------------------------------------
    def getBestTree(self):
        return self._best_tree

    def getParsimonyTree(self):
        """RAxML_parsimonyTree.<run> (RAxMLRunner.java:336-358): topology only; None when no parsimony run was made"""
        return getattr(self, "_parsimony_tree", None)

    def getParsimonyWithBLTree(self):
        """RAxML_result.<run>BL (RAxMLRunner.java:360-383); "" unless the parsimony-with-branch-lengths run was made"""
        return self._best_tree if self.algorithm == PARSIMONY_WITH_BL_ALGORITHM else ""

    def getPerSiteLLResultFile(self):
        return self._per_site_lines

    def getLikelihood(self):
        return self._lnl

    def getAlpha(self):
        return self._alpha

    def getSupportDecoratedTree(self, main_tree, support_trees):
        """RAxMLRunner.getSupportDecoratedTree (RAxMLRunner.java:453-516): `-f b`, integer percent labels"""
        try:
            return _e.support_tree(main_tree, list(support_trees), as_percent=True)
        except _e.EngineError as ex:
            self._fail(str(ex))
            return None

    def getBootstrapWeights(self, seed=None, reps=None):
        """replicate site-weight vectors of `-x seed -N reps` / `-b seed` (computeNextReplicate), pattern order"""
        aln = self._load()
        try:
            w, _ = aln.bootstrap_weights(self.random_seed if seed is None else seed, self.bootstrap_reps if reps is None else reps)
        finally:
            aln.close()
        return w

    def close(self):
        if self._own_ctx and self.ctx is not None:
            self.ctx.close()
            self.ctx = None


class B200FastTreeRunner:
    """drop-in for FastTreeRunner (FastTreeRunner.java:38-135), the tool PEPR calls ~100 times per refinement round for
    its support trees: `FastTree_WAG -gamma -nosupport X.faa` (newick on stdout), `-gamma` with local supports x 100
    truncated to integers when bootstrapReps > 0, optionally followed by raxmlHPC `-f e` branch lengths
    (getRaxmlBranchLengths, :142-199).  Here the tree comes from the engine's own ML search (parsimony start tree, lazy
    SPR, WAG+G4 branch lengths -- so it always carries `-f e` quality lengths), supports are replicate percentages.
    FastTree's CAT/ME heuristics are not reproduced: trees are compared by their splits (tests/test_gpu_search.py).
    Topological constraints (`-constraints file`, FastTreeRunner.java:53-83) are honoured: setConstraints takes FastTree's
    constraint alignment, setConstraintTree a tree (encoded exactly as getFastTreeConstraintsForTree does, :243-273); the start
    tree grows through constraint-preserving insertions only and the search scores only moves that keep every split.
    Refused loudly, not ignored: nucleotide alignments (`-gtr -nt`)."""

    def __init__(self, gpu=0, ctx=None, strict=False):
        self._ml = B200MLRunner(gpu=gpu, ctx=ctx, strict=strict)
        self.alignment, self.constraints, self.bootstrap_reps = None, None, 0
        self.use_raxml_branch_lengths, self.thread_count, self.nucleotide = False, 1, False
        self.result, self.last_error, self.strict = None, None, strict

    def setAlignment(self, alignment):
        self.alignment = alignment

    def setConstraints(self, constraints):
        self.constraints = constraints

    def setConstraintTree(self, tree_string):
        """FastTreeRunner.setConstraintTree (FastTreeRunner.java:224-229)"""
        if tree_string is not None:
            self.setConstraints(_e.constraints_from_tree(tree_string))

    def setBootstrapReps(self, n):
        self.bootstrap_reps = int(n)

    def setUseRaxmlBranchLengths(self, flag):
        self.use_raxml_branch_lengths = bool(flag)

    def setThreadCount(self, n):
        self.thread_count = int(n)

    def setNucleotide(self, flag):
        self.nucleotide = bool(flag)

    def run(self):
        self.result, self.last_error = None, None
        if self.nucleotide:
            self.last_error = "B200FastTreeRunner: nucleotide alignments are not supported by the engine"
            logger.error(self.last_error)
            if self.strict:
                raise _e.EngineError(self.last_error)
            return
        self._ml.setAlignment(self.alignment)
        self._ml.setBootstrapReps(self.bootstrap_reps)
        self._ml.constraints = self.constraints
        self._ml.algorithm = ML_ALGORITHM
        self._ml.start_tree = None
        self._ml.run()
        if self._ml.last_error is not None:
            self.last_error = self._ml.last_error   # reference convention: logged, result stays null
            return
        self.tree_options = "peprml search -m PROTGAMMAWAG" + ("" if self.bootstrap_reps else " -nosupport") + (
            " -constraints <%d columns>" % len(self.constraints.split("\n")[1]) if self.constraints else "")
        self.result = self._ml.getBestTreeWithSupports() if self.bootstrap_reps > 0 else self._ml.getBestTree()

    def getResult(self):
        return self.result

    def getLikelihood(self):
        return self._ml.getLikelihood()

    def close(self):
        self._ml.close()


# tree-building methods of PhylogeneticTreeBuilder (HandyConstants.java:36-68)
MAXIMUM_LIKELIHOOD, PARSIMONY, PARSIMONY_BL, FAST_TREE, NEIGHBOR_JOINING, MR_BAYES = "ml", "parsimony", "parsimony_bl", "FastTree", "nj", "mb"


class B200TreeBuilder:
    """drop-in for PhylogeneticTreeBuilder on the likelihood path (PhylogeneticTreeBuilder.java:97-129 run and the four build*
    methods :137-196): the method dispatch PEPR's pipeline talks to, with the engine's runners behind it.
      ml            -> B200MLRunner.run: getBestTree, or getBestTreeWithSupports when bootstrapReps > 0   (buildRaxmlTree)
      parsimony     -> setParsimonyOnly, getParsimonyTree                                                (buildRaxmlParsimonyTree)
      parsimony_bl  -> setParsimonyWithBL, getParsimonyWithBLTree                                        (buildRaxmlParsimonyTreeWithBL)
      FastTree      -> B200FastTreeRunner with constraint tree, bootstrap reps, raxml branch lengths     (buildFastTree)
    Neighbour joining and MrBayes are not on the accelerated path: run() records that in last_error and leaves the tree
    string None, which is what the reference's callers see when a tool fails (ExecUtilities.java:29-35)."""

    def __init__(self, gpu=0, ctx=None, strict=False):
        self._gpu, self._ctx, self.strict = gpu, ctx, strict
        self.alignment, self.tree_building_method, self.ml_matrix = None, MR_BAYES, "PROTGAMMAWAG"
        self.processes, self.bootstrap_reps = 1, 100          # PhylogeneticTreeBuilder.java:51-52
        self.use_raxml_branch_lengths, self.nucleotide, self.use_taxon_names = False, False, True
        self.constraint_tree, self.run_name, self.tree_string, self.last_error = None, None, None, None
        self.runner = None

    def setAlignment(self, alignment):
        self.alignment = alignment

    def getAlignment(self):
        return self.alignment

    def setTreeBuildingMethod(self, method):
        self.tree_building_method = method

    def setProcesses(self, n):
        self.processes = int(n)

    def setBootstrapReps(self, n):
        self.bootstrap_reps = int(n)

    def getBootstrapReps(self):
        return self.bootstrap_reps

    def setMLMatrix(self, m):
        self.ml_matrix = m

    def setConstraintTree(self, tree_string):
        self.constraint_tree = tree_string

    def setRunName(self, name):
        self.run_name = name

    def getRunName(self):
        return self.run_name

    def useRaxmlBranchLengths(self, flag):
        self.use_raxml_branch_lengths = bool(flag)

    def setNucleotide(self, flag):
        self.nucleotide = bool(flag)

    def getTreeString(self):
        return self.tree_string

    def setTreeString(self, s):
        self.tree_string = s

    def run(self):
        self.tree_string, self.last_error = None, None
        m = self.tree_building_method
        if m in (MAXIMUM_LIKELIHOOD, PARSIMONY, PARSIMONY_BL):
            r = self.runner = B200MLRunner(self.processes, gpu=self._gpu, ctx=self._ctx, strict=self.strict)
            try:
                r.setBootstrapReps(self.bootstrap_reps)
                r.setAlignment(self.alignment)
                if m == MAXIMUM_LIKELIHOOD:
                    r.setMatrix(self.ml_matrix)
                elif m == PARSIMONY:
                    r.setParsimonyOnly(True)
                else:
                    r.setParsimonyWithBL(True)
                r.run()
                self.last_error = r.last_error
                if r.last_error is None:
                    if m == MAXIMUM_LIKELIHOOD:
                        self.tree_string = r.getBestTree() if self.bootstrap_reps == 0 else r.getBestTreeWithSupports()
                    elif m == PARSIMONY:
                        self.tree_string = r.getParsimonyTree()
                    else:
                        self.tree_string = r.getParsimonyWithBLTree()
            finally:
                r.close()
        elif m == FAST_TREE:
            f = self.runner = B200FastTreeRunner(gpu=self._gpu, ctx=self._ctx, strict=self.strict)
            try:
                f.setNucleotide(self.nucleotide)
                f.setAlignment(self.alignment)
                f.setBootstrapReps(self.bootstrap_reps)
                f.setUseRaxmlBranchLengths(self.use_raxml_branch_lengths)
                if self.constraint_tree is not None:
                    f.setConstraintTree(self.constraint_tree)
                f.run()
                self.last_error, self.tree_string = f.last_error, f.getResult()
            finally:
                f.close()
        else:
            self.last_error = "B200TreeBuilder: method '%s' is not on the accelerated path (ml, parsimony, parsimony_bl, FastTree)" % m
            logger.error(self.last_error)
            if self.strict:
                raise _e.EngineError(self.last_error)


def gene_block_weights(block_lengths, keep, multiplicity=None):
    """Site weights of a gene-wise jackknife / bootstrap replicate over ONE concatenated alignment: the
    supermatrix of MSAConcatenator.concatenate (MSAConcatenator.java:78-189) lays the genes out one block after another;
    instead of concatenating half of the genes afresh for every support tree (PhylogenomicPipeline2.java:1227-1275) the
    replicate is the weight 1 on the columns of the genes in `keep` and 0 elsewhere (`multiplicity[g]` instead of 1 for a
    gene drawn several times).  -> int32 weight per column, for B200MLRunner.setSiteWeights / Alignment(site_weights=...)"""
    block_lengths = np.asarray(block_lengths, np.int64)
    w_gene = np.zeros(len(block_lengths), np.int32)
    for g in keep:
        w_gene[g] += 1 if multiplicity is None else int(multiplicity[g])
    return np.repeat(w_gene, block_lengths).astype(np.int32)


class JavaRandom:
    """java.util.Random (48-bit LCG), so that a seeded run draws the very sets the Java code would"""

    def __init__(self, seed):
        self.seed = (int(seed) ^ 0x5DEECE66D) & ((1 << 48) - 1)

    def _next(self, bits):
        self.seed = (self.seed * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        v = self.seed >> (48 - bits)
        return v - (1 << 32) if v >= (1 << 31) else v               # (int) cast of the Java original

    def nextIntAll(self):
        """Random.nextInt() without a bound"""
        return self._next(32)

    def nextInt(self, bound):
        if bound <= 0:
            raise ValueError("bound must be positive")
        if bound & (bound - 1) == 0:
            return (bound * self._next(31)) >> 31
        while True:
            bits = self._next(31)
            val = bits % bound
            if bits - val + (bound - 1) < (1 << 31):
                return val


def getRandomSet(length, min_value, max_value, allow_reuse, seed=None):
    """RandomSetUtils.getRandomSet (RandomSetUtils.java:9-35) with the seed the reference lacks (`new Random()` there makes
    every PEPR run draw different jackknife gene sets): same rejection loop over a java.util.Random stream; seed None draws a
    seed from the OS as the reference does.  -> list of ints, or None where the reference returns null"""
    rng_range = max_value - min_value + 1
    if not (allow_reuse or rng_range >= length):
        return None
    if seed is None:
        import os
        seed = int.from_bytes(os.urandom(6), "little")
    rnd = seed if isinstance(seed, JavaRandom) else JavaRandom(seed)
    used, out = [False] * rng_range, []
    for _ in range(length):
        pos = rnd.nextInt(rng_range)
        while not allow_reuse and used[pos]:
            pos = rnd.nextInt(rng_range)
        used[pos] = True
        out.append(min_value + pos)
    return out


def support_threads_for_memory(ntax, npatterns, free_bytes, requested):
    """The RAM throttle of buildConcatenatedTreeWithGeneWiseJackKnifeSupport (PhylogenomicPipeline2.java:1011-1036: `100 L N +
    2000 L` bytes per FastTree thread against the JVM's free RAM) restated for the engine: a replicate tree needs its own CLV
    arena of (ntax - 2) x patterns x 640 B (+ 4 B scaling counts and 8 B of parsimony scratch per node and pattern) in HBM, so
    that many replicate trees fit on one GPU side by side.  -> min(requested, what fits), at least 1 as in the reference"""
    per_tree = (ntax - 2) * npatterns * 644 + 2 * ntax * npatterns * 8 + ntax * npatterns
    return max(1, min(int(requested), int(free_bytes // max(per_tree, 1))))


class TreeSupportDecorator:
    """TreeSupportDecorator.addSupportValues (TreeSupportDecorator.java:86-163): raw counts as node labels"""

    @staticmethod
    def addSupportValues(main, supports):
        return _e.support_tree(main, list(supports), as_percent=False)


def getTreeScore(runner, tree):
    """PhylogenomicPipeline2.getTreeScore (PhylogenomicPipeline2.java:1482-1500): run `-f g` on one tree and sum line 2"""
    runner.setPerSiteLogLikelihoods(True)
    runner.setPerSiteLLTrees([tree])
    runner.run()
    lines = runner.getPerSiteLLResultFile()
    if not lines:
        return float("nan")
    return sum(float(x) for x in lines[1].split("\t")[1].split())
