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
