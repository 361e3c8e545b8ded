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

    def _load(self):
        a = self.alignment
        return _e.Alignment(self.ctx, a.names, a.seqs, alpha=1.0, model=self.matrix)

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
    Refused loudly, not ignored: topological constraints (`-constraints`) and nucleotide alignments (`-gtr -nt`)."""

    def __init__(self, gpu=0, ctx=None, strict=False):
        self._ml = B200MLRunner(gpu=gpu, ctx=ctx, strict=strict)
        self.alignment, self.constraints, self.bootstrap_reps = None, None, 0
        self.use_raxml_branch_lengths, self.thread_count, self.nucleotide = False, 1, False
        self.result, self.last_error, self.strict = None, None, strict

    def setAlignment(self, alignment):
        self.alignment = alignment

    def setConstraints(self, constraints):
        self.constraints = constraints

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
        if self.constraints is not None or self.nucleotide:
            self.last_error = "B200FastTreeRunner: constraints / nucleotide alignments are not supported by the engine"
            logger.error(self.last_error)
            if self.strict:
                raise _e.EngineError(self.last_error)
            return
        self._ml.setAlignment(self.alignment)
        self._ml.setBootstrapReps(self.bootstrap_reps)
        self._ml.algorithm = ML_ALGORITHM
        self._ml.start_tree = None
        self._ml.run()
        if self._ml.last_error is not None:
            self.last_error = self._ml.last_error   # reference convention: logged, result stays null
            return
        self.tree_options = "peprml search -m PROTGAMMAWAG" + ("" if self.bootstrap_reps else " -nosupport")
        self.result = self._ml.getBestTreeWithSupports() if self.bootstrap_reps > 0 else self._ml.getBestTree()

    def getResult(self):
        return self.result

    def getLikelihood(self):
        return self._ml.getLikelihood()

    def close(self):
        self._ml.close()


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
