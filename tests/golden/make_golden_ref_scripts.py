"""Record what the reference's own test scripts print on the all-reference stack (run in the build container only):

    python tests/golden/make_golden_ref_scripts.py      # writes tests/golden/ref_scripts/<script>.txt

tests/test_reference_scripts.py runs the same scripts, unmodified, on top of `install_as_punchclock()` and compares
every number they print.  SCRIPTS lists the reference tests that exercise the mirrored modules and run under the
stand-ins of _ref_stubs.py; a script that the stand-ins stop part-way (gym.make of the registered env id, which only
exists once the package __init__ has run) is recorded up to that point.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PUNCHCLOCK_REFERENCE", "/root/reference")
SCRIPTS = [
    "tests/common/test_visibility.py", "tests/common/test_agents.py", "tests/dynamics/test_dynamics.py",
    "tests/dynamics/test_dynamics_classes.py", "tests/dynamics/test_propagator.py", "tests/estimation/test_ez_ukf.py",
    "tests/estimation/test_ukf_v2.py", "tests/estimation/test_filter_base_class.py", "tests/environment/test_env.py",
    "tests/environment/test_env_utils.py", "tests/environment/test_env_parameters.py",
    "tests/schedule_tree/test_access_windows.py", "tests/simulation/test_sim_utils.py", "tests/common/test_metrics.py",
    "tests/common/test_utilities.py", "tests/common/test_orbits.py", "tests/policies/test_greedy_cov_v2.py",
    "tests/policies/test_random_policy.py", "tests/environment/test_info_wrappers.py",
    "tests/environment/test_misc_wrappers.py", "tests/environment/test_reward_wrappers.py",
]


def run(mode, script, ref=REF):
    # (PYTHONHASHSEED: the scripts print lists made from sets)
    r = subprocess.run([sys.executable, os.path.join(HERE, "_run_ref_script.py"), mode, ref, script],
                       capture_output=True, text=True, cwd="/tmp", timeout=900, env=dict(os.environ, PYTHONHASHSEED="0"))
    return r.returncode, r.stdout, r.stderr


def name_of(script):
    return script[len("tests/"):-3].replace("/", "_") + ".txt"


if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "ref_scripts"), exist_ok=True)
    for sc in SCRIPTS:
        rc, out, err = run("ref", sc)
        open(os.path.join(HERE, "ref_scripts", name_of(sc)), "w").write(out)
        print(f"{sc}: rc={rc}, {len(out.splitlines())} lines" + ("" if rc == 0 else "   stopped at: " + err.strip().splitlines()[-1][:120]))
