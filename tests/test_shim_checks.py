"""Host-side checks of the C++ operator shims that need no device."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_join_predicate_must_be_key_equality(tmp_path):
    """AlgHashJoinProbe / AlgNestJoinProbe evaluate joinpred_t::eval per visited node in the reference (algebra.hh:447,
    647-648); the device compares the hashed key attributes, so the shim self-checks the functor and throws otherwise."""
    exe = tmp_path / "pred_check"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-I", os.path.join(ROOT, "3d-hashjoin_b200", "hostcpp"),
                           os.path.join(ROOT, "tests", "cpp", "pred_check.cc"), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split() == ["0", "1", "1"], r.stdout + r.stderr


def test_ht_statistics_text_matches_the_reference_bytes(tmp_path):
    """HtStatistics::print / toCsvString / toCsvStringHeader of the shim (hostcpp/hj3d/ht_statistics.hh) against the bytes the
    unmodified reference writes for the same statistics (ht_statistics.cc:16-79; fixture from oracle/gen_stats_text_golden.py)."""
    import json
    exe = tmp_path / "stats_text"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-I", os.path.join(ROOT, "3d-hashjoin_b200", "hostcpp"),
                           os.path.join(ROOT, "tests", "cpp", "stats_text.cc"), "-o", str(exe)])
    names = ("num_buckets", "num_empty", "num_entries", "num_distinct_keys", "cc_min", "cc_max", "cc_sum", "cc_sumsq", "cc_count",
             "ccne_min", "ccne_max", "ccne_sum", "ccne_sumsq", "ccne_count")
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "ht_statistics_text.json")))["cases"]
    assert len(cases) >= 5
    for c in cases:
        out = subprocess.run([str(exe)] + [str(c["stats"][k]) for k in names], capture_output=True, text=True, check=True).stdout
        pr, csv, hdr = out.split("\x1e")
        assert (pr, csv, hdr) == (c["print"], c["csv"], c["header"]), c
