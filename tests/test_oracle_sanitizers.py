"""ASan / UBSan over the host oracle (the golden model is the checker of
everything else, so it gets checked itself).  CPU only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_golden_model_is_clean_under_asan_ubsan(tmp_path):
    exe = tmp_path / "golden_selftest"
    src = [os.path.join(ROOT, "oracle", "selftest.c"), os.path.join(ROOT, "oracle", "golden.c")]
    subprocess.run(["gcc", "-O1", "-g", "-std=gnu11", "-ffp-contract=off", "-fsanitize=address,undefined",
                    "-fno-sanitize-recover=all", "-pthread", *src, "-lm", "-o", str(exe)], check=True)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1")
    res = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    assert "SELFTEST_OK" in res.stdout
