"""Generate tests/golden/toy.npz from the reference's own fixture /root/reference/data/toy.rda.

Run in the build container (where /root/reference exists):

    python -m oracle.make_golden

The GPU box has no /root/reference, so tests read only the committed .npz.  Contents:
  coord (601,2), expression (500,601) [int8 -- the fixture is integer valued 0..44],
  gene_names, grid60 / grid100 (grid coordinates in the SCALED space, produced by
  singlecellhaystack_b200.synth.make_grid -- an input, not R's k-means),
  and the oracle's outputs for those grids: bandwidth, Q, nearest, D_KL (dense and sparse rows
  agree), plus the sparse randomization matrix and log10 p-values for grid100 with the
  hash-generated permutations / folds of singlecellhaystack_b200.synth (seed below).
Also writes tests/golden/toy_r1234.npz (see make_r_golden): the vignette call on R's own random stream.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import haystack_oracle as ho  # noqa: E402
from oracle.rda_reader import load_toy  # noqa: E402
from singlecellhaystack_b200 import synth  # noqa: E402

SEED = 20261019
REF_RDA = "/root/reference/data/toy.rda"


def main():
    toy = load_toy(REF_RDA)
    coord, expr = toy["coord"], toy["expression"]
    assert np.array_equal(expr, np.round(expr)) and expr.min() >= 0 and expr.max() <= 127
    xs, _, _ = ho.scale_coords(coord)
    out = dict(coord=coord, expression=expr.astype(np.int8), gene_names=np.array(toy["gene_names"]),
               seed=np.int64(SEED))
    for g in (60, 100):
        grid = synth.make_grid(xs, g, SEED)
        ds = ho.density_stage(xs, grid)
        dkl = ho.dkl_observed(expr, ds["density"], ds["Q"])
        dkl_sp = ho.dkl_observed(ho.RowSparse.from_dense(expr), ds["density"], ds["Q"])
        assert np.allclose(dkl, dkl_sp, rtol=1e-12, atol=0)
        out[f"grid{g}"] = grid
        out[f"bandwidth{g}"] = np.float64(ds["bandwidth"])
        out[f"Q{g}"] = ds["Q"]
        out[f"nearest{g}"] = ds["nearest"]
        out[f"dkl{g}"] = dkl
    # full path, sparse input, grid100
    grid = out["grid100"]
    res = ho.haystack_continuous_highD(coord, ho.RowSparse.from_dense(expr), grid, synth.perm_source(SEED),
                                       synth.cv_folds(SEED, 0, 100), synth.cv_folds(SEED, 1, 100))
    out["full_genes_to_randomize"] = res["genes_to_randomize"].astype(np.int32)
    out["full_KLD_rand"] = res["KLD_rand"]
    out["full_log_p"] = res["log_p_vals"]
    out["full_log_p_adj"] = res["log_p_adj"]
    out["full_best_df"] = np.array([res["fit"]["mean_model"]["best_df"], res["fit"]["sd_model"]["best_df"]])
    # full path, dense input (full permutations), grid100
    resd = ho.haystack_continuous_highD(coord, expr, grid, synth.perm_source(SEED),
                                        synth.cv_folds(SEED, 0, 100), synth.cv_folds(SEED, 1, 100))
    out["fulld_KLD_rand"] = resd["KLD_rand"]
    out["fulld_log_p"] = resd["log_p_vals"]
    dst = ROOT / "tests" / "golden"
    dst.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(dst / "toy.npz", **out)
    make_r_golden(coord, expr, toy["gene_names"], dst)
    top = np.argsort(res["log_p_vals"])[:10]
    print("top genes (oracle, synth grid):")
    for i in top:
        print(f"  {toy['gene_names'][i]:>9s}  D_KL {res['D_KL'][i]:.6f}  log.p {res['log_p_vals'][i]:.4f}")
    print("wrote", dst / "toy.npz", (dst / "toy.npz").stat().st_size, "bytes")


def make_r_golden(coord, expr, gene_names, dst):
    """tests/golden/toy_r1234.npz: the reference's vignette call `set.seed(1234); haystack(dat.tsne, dat.expression)`
    (docs/articles/a01_toy_example.html) replayed on the restated R generator / k-means (oracle/r_base): the grid R's
    kmeans() picks, the selected genes, the CV folds, a checksum per gene of the 100 x 601 permutation ids (the ids
    themselves are regenerated from the seed: 6 M integers are not a small fixture) and the oracle's results, which
    tests/test_r_parity_cpu.py holds against the table the reference printed."""
    from oracle import r_base as rb
    rng = rb.RRng(1234)
    xs, _, _ = ho.scale_coords(coord)
    grid = rb.get_grid_points(xs, rng, "centroid", 100)
    mean, sd = ho.row_mean_sd(expr)
    cv = (sd + 1e-300) / (mean + 1e-300)
    genes = ho.select_genes_to_randomize(cv, 100, "heavytails")
    perms = {int(g): rng.sample_int(601, 601, count=100) for g in genes}
    folds_mean = rng.sample(np.resize(np.arange(1, 11), 100))
    folds_sd = rng.sample(np.resize(np.arange(1, 11), 100))
    res = ho.haystack_continuous_highD(coord, expr, grid, lambda kind, gene, n, size, count: perms[gene],
                                       folds_mean, folds_sd)
    chk = np.array([int(perms[int(g)].astype(np.int64).dot(np.arange(1, 602)).sum() % (2 ** 31)) for g in genes])
    np.savez_compressed(dst / "toy_r1234.npz", seed=np.int64(1234), grid_scaled=grid,
                        grid_coordinates=res["grid_coordinates"], genes_to_randomize=genes.astype(np.int64),
                        folds_mean=folds_mean.astype(np.int32), folds_sd=folds_sd.astype(np.int32),
                        perm_checksums=chk, bandwidth=np.float64(res["bandwidth"]), Q=res["Q"], D_KL=res["D_KL"],
                        KLD_rand=res["KLD_rand"], log_p_vals=res["log_p_vals"], log_p_adj=res["log_p_adj"],
                        best_df=np.array([res["fit"]["mean_model"]["best_df"], res["fit"]["sd_model"]["best_df"]]))
    print("wrote", dst / "toy_r1234.npz", (dst / "toy_r1234.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
