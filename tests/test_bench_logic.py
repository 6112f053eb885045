"""CPU: the pieces of bench.py that decide what the JSON line claims -- the parity comparison and the choice of the committed ncu
capture behind `roofline.traffic` -- on synthetic inputs (no GPU, no reference run)."""
import json
import os

import numpy as np

import bench


def _logits(n, seed):
    return np.random.default_rng(seed).standard_normal((n, 64)).astype(np.float32)


def test_parity_identical_runs_report_bit_identity():
    ref = _logits(5, 0)
    toks = ["a", "b", "c", "d", "e"]
    par = bench.compare_parity(toks, ref, list(toks), ref.copy())
    assert par["greedy_ids_equal"] and par["first_divergence"] is None and par["tokens_compared"] == 5
    assert par["logits_bit_identical"] and par["logits_maxabs_over_range"] == 0.0 and par["logits_steps_compared"] == 5
    assert "reference's bits" in par["note"]


def test_parity_one_ulp_is_not_bit_identity():
    ref = _logits(4, 1)
    ours = ref.copy()
    ours[2, 7] = np.nextafter(ours[2, 7], np.float32(np.inf))
    par = bench.compare_parity(list("abcd"), ref, list("abcd"), ours)
    assert par["greedy_ids_equal"] and not par["logits_bit_identical"] and 0.0 < par["logits_maxabs_over_range"] < 1e-6


def test_parity_divergence_is_reported_with_the_reference_gap():
    ref = _logits(6, 2)
    ours = ref + np.float32(1e-3)
    par = bench.compare_parity(list("abcdef"), ref, list("abXdef"), ours)
    assert not par["greedy_ids_equal"] and par["first_divergence"] == 2
    assert par["logits_steps_compared"] == 3                     # steps 0 .. 2: the last one both arms evaluated on the same tokens
    assert "reference_top1_top2_gap_at_divergence" in par and not par["logits_bit_identical"]


def test_traffic_comes_from_the_newest_committed_capture():
    traffic, src = bench.committed_traffic()
    pdir = os.path.join(bench.ROOT, "profiles")
    newest = sorted(n for n in os.listdir(pdir) if n.endswith("_ncu_token_kernel.json"))[-1]
    cap = json.load(open(os.path.join(pdir, newest)))
    assert traffic == cap["dram_bytes_read"] + cap["dram_bytes_write"] and newest in src and "from_committed_profile" in src
    # the capture must be of the kernel the bench times: DRAM traffic within 2 % of the algorithmic bytes of a 7B q4_0 token
    assert abs(traffic / 4129423360 - 1.0) < 0.02
