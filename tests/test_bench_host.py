"""CPU tests of bench.py's host logic: the reference arm (oracle/_ref on a bounded sample), the cost model that
turns the sample's walls into a full-workload time, and the `config` both arms must share."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import ref_driver as rd  # noqa: E402

needs_ref = pytest.mark.skipif(not rd.available(), reason="oracle/_ref/libpawpy_ref.so not built")


def test_sample_plan_keeps_the_element_mix_and_site_lists_consistent():
    w = bench.workload("tiny")
    plan = bench.sample_plan(w, ns_each=2, nb=8, n_pair=4, kappa=1)
    R, S, cat = plan["R_sub"], plan["S_sub"], plan["cat"]
    assert w["vac"] in R and len(R) == len(S) + 1
    labR, labS = w["labels_R"][R], w["labels_S"][S]
    assert sorted(labS) == [0, 0, 1, 1]
    # matched pairs are the same physical site in both structures
    for mr, ms in zip(cat[0], cat[1]):
        assert np.allclose(w["coords_R"][R[mr]], w["coords_S"][S[ms]])
        assert labR[mr] == labS[ms]
    assert [R[i] for i in cat[2]] == [w["vac"]]
    full = bench.sample_plan(w, ns_each=10 ** 6, nb=8, n_pair=8, kappa=0)
    assert len(full["R_sub"]) == len(w["labels_R"]) and len(full["S_sub"]) == len(w["labels_S"])


def test_blocks_debug_option_keeps_the_first_blocks_and_says_so():
    """`--blocks n` (what one rank of an n-block job holds, for single-GPU traces of the e2e pipeline): the workload
    shrinks to the first k-points / one spin and its name says so, so such a line cannot pass for the real config."""
    w = bench.workload("tiny", blocks=1)
    assert w["nk"] == 1 and w["nspin"] == 1 and "[debug: first 1 block(s)]" in w["name"]
    assert bench.config_dict(w, 1, "strong")["kappa_blocks"] == 1
    full = bench.workload("tiny")
    assert full["nk"] * full["nspin"] == 4 and "debug" not in full["name"]
    assert np.array_equal(w["gvecs"][0], full["gvecs"][0])          # block 0 is the same block
    # level-1 / level-2 choice follows the block count
    assert bench.shard_mode(full, 4) == "kappa" and bench.shard_mode(w, 2) == "bands"


def test_config_dict_is_identical_for_both_arms():
    w = bench.workload("tiny")
    a = bench.config_dict(w, 1, "strong")
    b = bench.config_dict(bench.workload("tiny"), 1, "strong")
    assert a == b and a["kappa_blocks"] == 4 and a["pairs_per_step"] == 24 * 24 * 4


def test_ref_model_recovers_a_synthetic_cost_law():
    """Feed the model walls generated from known coefficients; it must reproduce the full-workload time."""
    w = bench.workload("tiny")
    plan = bench.sample_plan(w, ns_each=2, nb=16, n_pair=4, kappa=0)
    nsR, nsS, nb_s, n_pair = len(plan["R_sub"]), len(plan["S_sub"]), plan["nb"], plan["n_pair"]
    c_site, c_f, c_ps, c_dot, c_cp, a_ps, a_cp = 0.05, 2e-3, 4e-5, 7e-5, 3e-7, 1e-4, 2e-5
    t = {"site_R": c_site * nsR, "site_1": c_site,
         "setup_R": c_site * nsR + nb_s * (c_f + c_ps * nsR), "setup_S": c_site * nsS + nb_s * (c_f + c_ps * nsS),
         "setup_1site": c_site + nb_s * (c_f + c_ps),
         "overlap_setup": c_site + nb_s * (c_f + c_ps),
         "pseudo": n_pair * (a_ps + nb_s * c_dot), "pseudo_8band_basis": n_pair * (a_ps + 8 * c_dot),
         "compensation": nb_s * (a_cp + nb_s * c_cp * nsS), "compensation_no_sites": nb_s * a_cp}
    m = bench.ref_model(w, plan, t, threads=1)
    NR, NS, nb, NK = len(w["labels_R"]), len(w["labels_S"]), w["nband"], 4
    want = (c_site * (NR + NS) + NK * nb * (2 * c_f + c_ps * (NR + NS)) + c_site + NK * nb * (c_f + c_ps)
            + NK * nb * (a_ps + nb * c_dot) + NK * nb * (a_cp + nb * c_cp * NS))
    assert abs(m["full_workload_s"] - want) / want < 1e-9
    assert abs(m["pairs_per_s"] - nb * nb * NK / want) / m["pairs_per_s"] < 1e-9


@needs_ref
def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` on the tiny config: one JSON line, same config keys as the B200 arm, e2e block
    with zero copies, ms_per_step = the measured sample step (so steps x ms_per_step fits the run)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "tiny",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 1 and line["scaling"] == "strong"
    assert line["config"] == bench.config_dict(bench.workload("tiny"), 1, "strong")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["steps"] * line["ms_per_step"] * 1e-3 <= line["wall_s"]
    assert line["value"] > 0 and all(v >= 0 for v in line["cpu_baseline"]["model"]["stages_full_s"].values())
