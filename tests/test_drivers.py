"""The drop-in C++ drivers (3d-hashjoin_b200/drivers, written against hostcpp/hj3d/algebra.hh = the reference's
operator surface): generated relations, CSV schema and -- on the GPU -- every counter the reference's own
drivers write for the same command line (golden fixtures from the unmodified reference)."""
import csv
import json
import os
import subprocess

import numpy as np
import pytest

from helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "3d-hashjoin_b200", "drivers")

EXP1_HEADER = ("mintime;minreps;log2CardR;log2CardS;skew;t;fkMax;numDvSa;b;plan;ht_impl;build;probe;ht_buckets;ht_fracEmpty;"
               "cc0_avg;cc0_min;cc0_max;cc1_avg;cc1_min;cc1_max;reps;t_total;t_buildStr;t_probeStr;t_top;c_scanBuild;c_selBuild;"
               "c_htBuild;c_scanProbe;c_selProbe;c_htProbe;c_htProbeCmp;c_unnest;c_top")           # main_experiment1.cc:1289-1333
EXP4_HEADER = ("mintime;minreps;log2CardR;a;aM;b;bM;cardR;cardS;cardT;plan;ht_impl;reps;t_total;t_build_S;t_build_T;t_probe_R;"
               "c_sc_R;c_sc_S;c_sc_T;c_build_S;c_build_T;c_probe_RS;c_probe_RS_cmp;c_probe_RT;c_probe_RT_cmp;c_unnest_S;"
               "c_unnest_T;c_top")                                                                    # main_experiment4.cc:770-812


def build_drivers():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "3d-hashjoin_b200"), "drivers"])


def dump(*args):
    out = subprocess.run([os.path.join(DRV, "datagen_dump.out"), *map(str, args)], capture_output=True, text=True, check=True).stdout
    return out.split("\n")


@pytest.mark.parametrize("name,args", [("exp1_R10_S12_uni_t0_b1", (10, 12, 0, 0)), ("exp1_R12_S14_zipf_t2_b2", (12, 14, 1, 2)),
                                       ("exp1_R8_S13_zipf_t0_b4", (8, 13, 1, 0))])
def test_exp1_generator_equals_reference(name, args):
    """hj3d/datagen.hh reproduces Experiment1::init bit for bit (uniform and Zipf, libstdc++)."""
    build_drivers()
    z, meta = load_golden(name)
    o = dump("exp1", *args)
    assert int(o[0]) == meta["numDvSa"]
    assert np.array_equal(np.array(o[1].split(), dtype=np.uint32), z["Rk"])
    assert np.array_equal(np.array(o[2].split(), dtype=np.uint32), z["Sa"])


@pytest.mark.parametrize("name,args", [("exp4_R12_a4_b3_A5_B7", (12, 4, 5, 3, 7)), ("exp4_R10_a2_b2_A10_B1", (10, 2, 10, 2, 1))])
def test_exp4_generator_equals_reference(name, args):
    build_drivers()
    z, _ = load_golden(name)
    o = dump("exp4", *args)
    assert np.array_equal(np.array(o[0].split(), dtype=np.uint32), z["Sa"])
    assert np.array_equal(np.array(o[1].split(), dtype=np.uint32), z["Ta"])


def test_drivers_fail_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    build_drivers()
    r = subprocess.run([os.path.join(DRV, "main_experiment1.out"), "-R", "8", "-S", "10", "--no-skew", "-t", "0",
                        "--measure-file", str(tmp_path / "m.csv"), "-p", "Csr"], capture_output=True, text=True)
    assert r.returncode != 0 and "hj3d" in r.stderr          # no CPU fallback behind the operator templates


def test_sharded_example_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    build_drivers()
    r = subprocess.run([os.path.join(DRV, "main_sharded_example.out"), "-R", "8", "-S", "10", "-g", "2"], capture_output=True, text=True)
    assert r.returncode == 3 and "hj3d" in r.stderr and "CPU fallback" in r.stderr     # hj3d_ctx_create's own message


def read_csv(path):
    rows = list(csv.reader(open(path), delimiter=";"))
    return rows[0], rows[1:]


@pytest.mark.gpu
@pytest.mark.parametrize("name,cli", [("exp1_R10_S12_uni_t0_b1", ["-R", "10", "-S", "12", "--no-skew", "-t", "0", "-b", "1"]),
                                      ("exp1_R12_S14_zipf_t2_b2", ["-R", "12", "-S", "14", "--skew", "-t", "2", "-b", "2"])])
def test_experiment1_driver_csv_matches_reference(tmp_path, name, cli):
    build_drivers()
    _, meta = load_golden(name)
    f = tmp_path / "exp1.csv"
    r = subprocess.run([os.path.join(DRV, "main_experiment1.out"), *cli, "--measure-file", str(f)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    hdr, rows = read_csv(f)
    assert ";".join(hdr) == EXP1_HEADER
    by_plan = {}
    for row in rows:
        plan = row[9]
        if plan in ("scr", "scs"):
            assert len(row) == len(hdr) - 1                  # reference quirk: no `reps` field
            continue
        assert len(row) == len(hdr)
        by_plan[plan] = dict(zip(hdr, row))
    assert set(by_plan) == {"Csr", "CsrUU", "Crs", "Nsr", "Nrs", "NrsNU"}
    nR, nS = 1 << meta["log2R"], 1 << meta["log2S"]
    for plan, g in meta["plans"].items():
        row, st = by_plan[plan], g["stats"]
        assert int(row["numDvSa"]) == meta["numDvSa"] and int(row["fkMax"]) == meta["fkMax"]
        assert int(row["ht_buckets"]) == g["D"]
        assert abs(float(row["ht_fracEmpty"]) - st["num_empty"] / st["num_buckets"]) < 1e-5
        assert abs(float(row["cc0_avg"]) - st["cc_sum"] / st["cc_count"]) < 1e-4
        assert (int(row["cc0_min"]), int(row["cc0_max"])) == (st["cc_min"], st["cc_max"])
        assert abs(float(row["cc1_avg"]) - st["ccne_sum"] / st["ccne_count"]) < 1e-4
        assert (int(row["cc1_min"]), int(row["cc1_max"])) == (st["ccne_min"], st["ccne_max"])
        assert int(row["c_htProbe"]) == g["probe"]["matches"]
        assert int(row["c_htProbeCmp"]) == g["probe"]["num_cmps"]
        build_on_R = plan in ("Csr", "CsrUU", "Nsr")
        assert int(row["c_htBuild"]) == (nR if build_on_R else nS) and int(row["c_scanProbe"]) == (nS if build_on_R else nR)
        if plan in ("Nsr", "Nrs"):
            assert int(row["c_unnest"]) == g["unnest"]["out_tuples"] == nS
        else:
            assert row["c_unnest"] == "NA"
        assert int(row["c_top"]) == (meta["numDvSa"] if plan == "NrsNU" else nS)
        assert int(row["reps"]) >= 8


@pytest.mark.gpu
def test_experiment4_driver_csv_matches_reference(tmp_path):
    build_drivers()
    _, meta = load_golden("exp4_R12_a4_b3_A5_B7")
    f = tmp_path / "exp4.csv"
    r = subprocess.run([os.path.join(DRV, "main_experiment4.out"), "-R", "12", "-a", "4", "-b", "3", "-A", "5", "-B", "7",
                        "--measure-file", str(f)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    hdr, rows = read_csv(f)
    assert ";".join(hdr) == EXP4_HEADER
    by_plan = {row[10]: dict(zip(hdr, row)) for row in rows}
    for plan in ("Ndu", "Chj"):
        row, g = by_plan[plan], meta[plan]
        for col, key in (("c_probe_RS", "c_probe_RS"), ("c_probe_RS_cmp", "c_probe_RS_cmp"), ("c_probe_RT", "c_probe_RT"),
                         ("c_probe_RT_cmp", "c_probe_RT_cmp"), ("c_top", "c_top")):
            assert int(row[col]) == g[key], (plan, col)
    assert (int(by_plan["Ndu"]["c_unnest_S"]), int(by_plan["Ndu"]["c_unnest_T"])) == (meta["Ndu"]["c_unnest1"], meta["Ndu"]["c_unnest2"])
    assert by_plan["Chj"]["c_unnest_S"] == "NA"


@pytest.mark.gpu
def test_algebra_example_prints_reference_results_in_reference_order():
    """main_algebra_example: the printed result rows and operator counts of algebra_test0..3."""
    build_drivers()
    r = subprocess.run([os.path.join(DRV, "main_algebra_example.out")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    t = {k: out.split(k)[1].split("test")[0] for k in ("test0", "test1", "test2", "test3")}
    assert "(1,11)\n(2,21)\n(3,31)\n" in t["test0"] and "AlgTop|3|" in t["test0"] and "AlgScan|4|" in t["test0"]
    assert "(1,11,1,-1)\n(2,21,2,-1)\n(3,31,3,-1)\n" in t["test1"] and "AlgNestJoinBuild|6|" in t["test1"] and "AlgNestJoinProbe|3|" in t["test1"]
    joined = "(1,11,1,-1)\n(1,11,1,-3)\n(1,11,1,-2)\n(2,21,2,-1)\n(2,21,2,-2)\n(3,31,3,-1)\n"
    assert joined in t["test2"] and "AlgUnnest|6|" in t["test2"] and "AlgTop|6|" in t["test2"] and "AlgNestJoinProbe|3|" in t["test2"]
    assert joined in t["test3"] and "AlgHashJoinProbe|6|" in t["test3"] and "sizeof(Node) 24" in t["test3"]


@pytest.mark.gpu
@pytest.mark.parametrize("cli", [["-R", "14", "-S", "18", "-g", "3", "-p", "Csr"], ["-R", "14", "-S", "17", "-g", "2", "-p", "Nsr"],
                                 ["-R", "14", "-S", "18", "-g", "4", "-p", "Csr", "--skew"], ["-R", "13", "-S", "18", "-g", "2", "-p", "Nsr", "--skew"]])
def test_sharded_example_driver(cli):
    """drivers/main_sharded_example.cc: a C++ host that owns all ranks in one process, C ABI only: host relations streamed
    through the exchange (uniform) or hot-key probe replication (--skew); merged counters, checksum and HtStatistics must
    equal the unsharded hj3d_join_host result (the driver exits non-zero otherwise)."""
    build_drivers()
    r = subprocess.run([os.path.join(DRV, "main_sharded_example.out"), *cli], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    hdr, row = [l.split(";") for l in r.stdout.strip().split("\n")[-2:]]
    rec = dict(zip(hdr, row))
    assert rec["sharded_equals_unsharded"] == "yes" and int(rec["c_top"]) == 1 << int(cli[3])
    if "--skew" in cli:
        assert int(rec["hot_tuples_kept_local"]) > 0

