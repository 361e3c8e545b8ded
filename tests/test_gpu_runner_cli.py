"""The reference-facing layers on the GPU: the RAxMLRunner mirror (in-process) and the raxmlHPC-compatible executable
(process + files, the protocol PEPR's ExecUtilities uses)."""
import os
import re
import subprocess

import numpy as np
import pytest

import pepr_b200 as pb
from oracle import oracle as orc
from pepr_b200 import runner as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "pepr_b200", "bin", "peprml")


def test_runner_branch_lengths_like_f_e(gpu_ctx, golden):
    g = golden("small")
    r = R.B200MLRunner(threads=4, ctx=gpu_ctx)
    r.setAlignment(R.SequenceAlignment(g.names, g.seqs))
    r.setMatrix("PROTGAMMAWAG")
    r.setStartTree(g.meta["tree_in"])
    r.run()
    assert r.last_error is None
    assert abs(r.getLikelihood() - g.meta["fe"]["lnl"]) < 0.1          # raxmlHPC's own stopping tolerance
    assert abs(r.getAlpha() - g.meta["fe"]["alpha"]) / g.meta["fe"]["alpha"] < 0.05
    t = r.getBestTree()
    assert t.endswith("):0.0;") and all(n in t for n in g.names)


def test_runner_per_site_ll_and_tree_score(gpu_ctx, golden):
    g = golden("small")
    r = R.B200MLRunner(ctx=gpu_ctx)
    r.setAlignment(R.SequenceAlignment(g.names, g.seqs))
    score = R.getTreeScore(r, g.meta["tree_in"])
    lines = r.getPerSiteLLResultFile()
    assert lines[0].split() == ["1", "300"]
    vals = np.array([float(x) for x in lines[1].split("\t")[1].split()])
    assert len(vals) == 300
    # raxmlHPC -f g optimises first too; its per-site values are the golden ones up to the optimisers' 0.1 lnL agreement
    assert abs(score - sum(g.meta["fg"]["per_site"])) < 0.1
    assert np.abs(vals - np.array(g.meta["fg"]["per_site"])).max() < 0.05


def test_runner_reports_errors_the_reference_way(gpu_ctx, golden):
    g = golden("small")
    r = R.B200MLRunner(ctx=gpu_ctx)
    r.setAlignment(R.SequenceAlignment(g.names, g.seqs))
    r.setStartTree("((TaxA,TaxB),(TaxC,Oops),((TaxE,TaxF),(TaxG,TaxH)));")
    r.run()
    assert r.getBestTree() == "" and "Oops" in r.last_error
    r.setMatrix("GTRGAMMA")                     # PEPR only ever asks for PROTGAMMAWAG; anything else is an error, not a guess
    r.run()
    assert r.getBestTree() == "" and "PROTGAMMAWAG" in r.last_error
    with pytest.raises(pb.EngineError):
        R.B200MLRunner(ctx=gpu_ctx, strict=True).run()


def test_runner_bootstrap_weights_and_supports(gpu_ctx, golden):
    g = golden("small")
    r = R.B200MLRunner(ctx=gpu_ctx)
    r.setAlignment(R.SequenceAlignment(g.names, g.seqs))
    r.setBootstrapReps(3)
    assert r.getBootstrapWeights(seed=g.meta["fj"]["seed"]).tolist() == g.meta["fj"]["replicate_weights"]
    from tests.test_oracle_golden import _labels
    got = r.getSupportDecoratedTree(g.meta["fe"]["tree"], g.meta["fb"]["support_trees"])
    assert _labels(got) == _labels(g.meta["fb"]["bipartitions"])


def _run_cli(args, cwd, env=None):
    return subprocess.run([CLI] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                          env=dict(os.environ, **(env or {})))


def test_cli_f_e_writes_raxml_files(tmp_path, golden):
    g = golden("small")
    from pepr_b200 import synth
    synth.write_phylip(str(tmp_path / "t.phy"), g.names, g.seqs)
    (tmp_path / "t.nwk").write_text(g.meta["tree_in"] + "\n")
    r = _run_cli(["-f", "e", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "run1", "-t", "t.nwk", "-T", "8"], tmp_path)
    assert r.returncode == 0, r.stderr
    info = (tmp_path / "RAxML_info.run1").read_text()
    lnl = float(re.search(r"Final GAMMA\s+likelihood: (\S+)", info).group(1))
    alpha = float(re.search(r"alpha: (\S+)", info).group(1))
    assert abs(lnl - g.meta["fe"]["lnl"]) < 0.1
    assert abs(alpha - g.meta["fe"]["alpha"]) / g.meta["fe"]["alpha"] < 0.05
    assert "Alignment has %d distinct alignment patterns" % g.meta["fe"]["patterns"] in info
    tree = (tmp_path / "RAxML_result.run1").read_text().strip()
    assert tree.endswith(":0.0;")
    # like raxmlHPC, a second run under the same name is refused
    assert _run_cli(["-f", "e", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "run1", "-t", "t.nwk"], tmp_path).returncode != 0


def test_cli_f_g_and_f_j(tmp_path, golden):
    g = golden("small")
    from pepr_b200 import synth
    from oracle import oracle as orc
    synth.write_phylip(str(tmp_path / "t.phy"), g.names, g.seqs)
    (tmp_path / "t.nwk").write_text(g.meta["tree_in"] + "\n")
    r = _run_cli(["-f", "g", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "g1", "-z", "t.nwk"], tmp_path)
    assert r.returncode == 0, r.stderr
    lines = (tmp_path / "RAxML_perSiteLLs.g1").read_text().split("\n")
    assert lines[0].split() == ["1", "300"]
    vals = [float(x) for x in lines[1].split("\t")[1].split()]
    assert abs(sum(vals) - sum(g.meta["fg"]["per_site"])) < 0.1
    r = _run_cli(["-f", "j", "-b", "12345", "-#", "3", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "j1"], tmp_path)
    assert r.returncode == 0, r.stderr
    for k in range(3):
        _, bs = orc.read_phylip(str(tmp_path / ("t.phy.BS%d" % k)))
        bc = orc.encode(bs)
        cols = {}
        for j in range(bc.shape[1]):
            key = bytes(bc[:, j])
            cols[key] = cols.get(key, 0) + 1
        w = [cols.get(bytes(g.pat[:, p]), 0) for p in range(g.pat.shape[1])]
        assert w == g.meta["fj"]["replicate_weights"][k]         # bit exact with raxmlHPC -f j


def test_cli_rejects_unknown_model_loudly(tmp_path, golden):
    g = golden("small")
    from pepr_b200 import synth
    synth.write_phylip(str(tmp_path / "t.phy"), g.names, g.seqs)
    (tmp_path / "t.nwk").write_text(g.meta["tree_in"] + "\n")
    r = _run_cli(["-f", "e", "-m", "GTRGAMMA", "-s", "t.phy", "-n", "m1", "-t", "t.nwk"], tmp_path)
    assert r.returncode != 0 and "PROTGAMMAWAG" in r.stderr


def test_cli_f_d_and_f_a_tree_search(tmp_path, golden):
    """PEPR's default full-tree call (`-f d`) and the rapid-bootstrap call (`-f a -x seed -N reps`) through the executable"""
    import gzip
    from tests.test_gpu_search import _splits, _strip
    g = golden("search")
    from pepr_b200 import synth
    synth.write_phylip(str(tmp_path / "t.phy"), g.names, g.seqs)
    r = _run_cli(["-f", "d", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "d1", "-T", "4"], tmp_path)
    assert r.returncode == 0, r.stderr
    info = (tmp_path / "RAxML_info.d1").read_text()
    lnl = float(re.search(r"Final GAMMA-based Score of best tree (\S+)", info).group(1))
    assert lnl >= g.meta["fd"]["lnl"] - 0.5
    tree = (tmp_path / "RAxML_result.d1").read_text().strip()
    assert _splits(_strip(tree)) == _splits(g.meta["fd"]["tree"])
    assert (tmp_path / "RAxML_bestTree.d1").exists()
    r = _run_cli(["-f", "a", "-x", "12345", "-N", "5", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "a1"], tmp_path)
    assert r.returncode == 0, r.stderr
    assert len((tmp_path / "RAxML_bootstrap.a1").read_text().strip().split("\n")) == 5
    bip = (tmp_path / "RAxML_bipartitions.a1").read_text().strip()
    labels = [int(x) for x in re.findall(r"\)(\d+)", bip)]
    assert len(labels) == len(g.names) - 3 and min(labels) >= 0 and max(labels) <= 100
    r = _run_cli(["-f", "d", "-y", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "p1"], tmp_path)     # parsimony tree only
    assert r.returncode == 0 and (tmp_path / "RAxML_parsimonyTree.p1").exists()


def test_concurrent_contexts_from_threads(golden):
    """PEPR runs up to `tree_threads` runners concurrently in one JVM (PhylogenomicPipeline2.java:1233-1254): distinct contexts
    on the same GPU used from distinct host threads at the same time must give the single-threaded answers"""
    import threading
    cases = ["small", "dup", "deep", "aquificales"]
    want, got, errs = {}, {}, []

    def work(case, out):
        try:
            g = golden_get(case)
            ctx = pb.Context(0)
            aln = pb.Alignment(ctx, g.names, g.seqs, alpha=1.0)
            tree = pb.Tree(aln, g.meta["tree_in"])
            out[case] = tree.optimize(True, 0.1) + (tree.newick(),)
            tree.close(); aln.close(); ctx.close()
        except Exception as e:  # noqa: BLE001
            errs.append((case, repr(e)))

    from tests.conftest import Golden
    cache = {c: Golden(c) for c in cases}
    golden_get = cache.__getitem__
    for c in cases:
        work(c, want)
    threads = [threading.Thread(target=work, args=(c, got)) for c in cases]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for c in cases:
        assert got[c] == want[c], c       # bit-identical: same kernels, same fixed-order reductions


def test_cli_on_the_reference_example_genomes(tmp_path, golden):
    """the executable on the real-data supermatrix (examples/Aquificales, tests/golden/make_real.py): the files PEPR reads
    back after `-f e` and `-f g` carry raxmlHPC's numbers; long taxon names survive the relaxed phylip round trip"""
    g = golden("aquificales")
    from pepr_b200 import synth
    synth.write_phylip(str(tmp_path / "a.phy"), g.names, g.seqs)
    (tmp_path / "a.nwk").write_text(g.meta["tree_in"] + "\n")
    r = _run_cli(["-f", "e", "-m", "PROTGAMMAWAG", "-s", "a.phy", "-n", "fe", "-t", "a.nwk"], tmp_path)
    assert r.returncode == 0, r.stderr
    info = (tmp_path / "RAxML_info.fe").read_text()
    lnl = float(re.search(r"Final GAMMA\s+likelihood: (\S+)", info).group(1))
    assert abs(lnl - g.meta["fe"]["lnl"]) < 0.1
    assert "Alignment has %d distinct alignment patterns" % g.meta["fe"]["patterns"] in info
    tree = (tmp_path / "RAxML_result.fe").read_text().strip()
    assert all(n in tree for n in g.names)
    (tmp_path / "ref.nwk").write_text(g.meta["fe"]["tree"] + "\n")
    r = _run_cli(["-f", "g", "-m", "PROTGAMMAWAG", "-s", "a.phy", "-n", "fg", "-z", "ref.nwk"], tmp_path)
    assert r.returncode == 0, r.stderr
    lines = (tmp_path / "RAxML_perSiteLLs.fg").read_text().split("\n")
    got = np.array([float(x) for x in lines[1].split("\t")[1].split()])
    ref = np.array(g.meta["fg"]["per_site"])
    # `-f g` re-optimises the model on the given tree like raxmlHPC does, so alpha (hence every site) moves in the 4th decimal
    assert got.shape == ref.shape and abs(got.sum() - ref.sum()) < 0.1 and np.abs(got - ref).max() < 5e-3


def _splits(newick):
    sets = []
    taxa = sorted(orc._leafsets(orc._parse_topology(newick), sets))
    return {orc._canon(x, taxa) for x in sets if 1 < len(x) < len(taxa) - 1}


def test_runner_parsimony_with_branch_lengths_is_y_then_f_e(gpu_ctx, golden):
    """runRaxmlParsimonyWithBranchLengths (RAxMLRunner.java:215-280) = `-f d -y`, then `-f e -t` on that topology -- NOT a
    topology search.  Golden `fy`: what raxmlHPC gives for exactly that pair of runs on the `search` alignment."""
    g = golden("search")
    fy = g.meta["fy"]
    r = R.B200MLRunner(ctx=gpu_ctx)
    r.setAlignment(R.SequenceAlignment(g.names, g.seqs))
    # (1) the `-f e` half on raxmlHPC's own parsimony tree: its lnL within raxmlHPC's stopping tolerance
    r.setStartTree(fy["parsimony_tree"])
    r.run()
    assert r.last_error is None and abs(r.getLikelihood() - fy["lnl"]) < 0.1, (r.getLikelihood(), fy["lnl"])
    assert r.getParsimonyWithBLTree() == ""                     # not a parsimony-with-BL run: the reference finds no such file
    # (2) the whole thing: the engine's own parsimony tree (another random addition order), branch lengths on THAT topology
    r2 = R.B200MLRunner(ctx=gpu_ctx)
    r2.setAlignment(R.SequenceAlignment(g.names, g.seqs))
    r2.setParsimonyWithBL(True)
    r2.run()
    assert r2.last_error is None
    ptree, bl = r2.getParsimonyTree(), r2.getParsimonyWithBLTree()
    assert ":" not in ptree and bl.endswith("):0.0;")
    assert _splits(bl.replace("):0.0;", ");")) == _splits(ptree)            # the topology was not searched
    host, _ = pb.parsimony_tree(g.names, g.seqs, 12345)
    assert _splits(host.replace("):0.0;", ");")) == _splits(ptree)          # it IS the stepwise-addition parsimony tree
    # two near-optimal parsimony trees of one alignment: likelihoods close, both below the ML search's tree
    assert abs(r2.getLikelihood() - fy["lnl"]) < 40.0 and r2.getLikelihood() <= g.meta["fd"]["lnl"] + 0.5
    # parsimony only (`-y`): no likelihood at all
    r3 = R.B200MLRunner(ctx=gpu_ctx)
    r3.setAlignment(R.SequenceAlignment(g.names, g.seqs))
    r3.setParsimonyOnly(True)
    r3.run()
    assert r3.getParsimonyTree() == ptree and r3.getBestTree() == "" and r3.getLikelihood() is None


def test_cli_spreads_one_call_over_several_gpus(tmp_path, golden):
    """PEPRML_GPUS=2: what `-T n` is for raxmlHPC-PTHREADS (RAxMLRunner.java:130-132) -- one call, the patterns sharded over a
    group of GPUs inside the one process, rank 0 writing the files.  Same lnL, alpha, tree and per-site lnL as the one-GPU
    call (the sums are added shard by shard, so agreement is to rounding, not to the bit), and the tree search (`-f d`,
    parsimony scans + lazy SPR in lock step) ends on the same tree."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    g = golden("wide")
    from pepr_b200 import synth
    synth.write_phylip(str(tmp_path / "w.phy"), g.names, g.seqs)
    (tmp_path / "w.nwk").write_text(g.meta["tree_in"] + "\n")
    out = {}
    for tag, env in (("one", {}), ("two", {"PEPRML_GPUS": "2"})):
        r = _run_cli(["-f", "g", "-m", "PROTGAMMAWAG", "-s", "w.phy", "-n", tag, "-z", "w.nwk", "-T", "2"], tmp_path, env=env)
        assert r.returncode == 0, r.stderr
        info = (tmp_path / ("RAxML_info." + tag)).read_text()
        lnl = float(re.search(r"Final GAMMA\s+likelihood: (\S+)", info).group(1))
        alpha = float(re.search(r"alpha: (\S+)", info).group(1))
        ps = np.array([float(x) for x in (tmp_path / ("RAxML_perSiteLLs." + tag)).read_text().split("\n")[1].split("\t")[1].split()])
        out[tag] = (lnl, alpha, ps, (tmp_path / ("RAxML_result." + tag)).read_text().strip())
    a, b = out["one"], out["two"]
    assert abs(a[0] - b[0]) < 1e-3 and abs(a[1] - b[1]) < 1e-4 and np.abs(a[2] - b[2]).max() < 1e-4
    la = [float(x) for x in re.findall(r":([0-9.]+)", a[3])]
    lb = [float(x) for x in re.findall(r":([0-9.]+)", b[3])]
    assert len(la) == len(lb) and np.allclose(la, lb, atol=1e-5)
    s = golden("search")
    synth.write_phylip(str(tmp_path / "s.phy"), s.names, s.seqs)
    trees = {}
    for tag, env in (("d1", {}), ("d2", {"PEPRML_GPUS": "2"})):
        r = _run_cli(["-f", "d", "-m", "PROTGAMMAWAG", "-s", "s.phy", "-n", tag, "-p", "12345"], tmp_path, env=env)
        assert r.returncode == 0, r.stderr
        trees[tag] = re.sub(r":[0-9.eE+-]+", "", (tmp_path / ("RAxML_bestTree." + tag)).read_text().strip())
    assert trees["d1"] == trees["d2"]
