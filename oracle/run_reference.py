"""Run the UNMODIFIED upstream pore scripts for K timesteps and dump intermediate state.

Test infrastructure, build container only (needs /root/reference).  The script text is read from
the reference tree at run time, patched in memory in exactly two ways -- the loop bound
(`range(num_timesteps)` -> `range(K)`) and three calls to a dump hook -- and
written to a scratch directory outside the repository, where it runs as `__main__` with the
matplotlib stub on PYTHONPATH.  Nothing of the reference is copied into the repo.

usage: python oracle/run_reference.py {pore|temp} K OUTDIR
Dumps OUTDIR/{walls,end}_{step}.npz (full particle state after the wall phase incl. recapture and
at the end of the step) and OUTDIR/final.npz (completed path lists, per-step series).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AMC_REFERENCE_DIR", "/root/reference")

HOOK = '''
import os as _os
def _amc_dump(tag, step):
    np.savez(_os.path.join(_os.environ["AMC_DUMP_DIR"], "%s_%d.npz" % (tag, step)),
             x=x_vals, y=y_vals, z=z_vals, vx=x_velocities, vy=y_velocities, vz=z_velocities,
             dist=dist_since_collision, dist_x=dist_x_since_collision, dist_y=dist_y_since_collision,
             dist_z=dist_z_since_collision, flag=full_path_traveled,
             ncol=np.int64(num_collisions_per_step.value))
def _amc_final(**kw):
    np.savez(_os.path.join(_os.environ["AMC_DUMP_DIR"], "final.npz"), **{k: np.array([float(v) for v in vs]) for k, vs in kw.items()})
'''


def patched_source(kind, k):
    name = {"pore": "Open_Air_Pore_MC.py", "temp": "Temperature_Pore_MC.py"}[kind]
    src = open(os.path.join(REF, name)).read()
    var = "i" if kind == "pore" else "step"
    assert src.count("range(num_timesteps)") == 1
    src = src.replace("range(num_timesteps)", "range(%d)" % k)
    a1 = "            # PARTICLE-PARTICLE COLLISIONS\n"
    a2 = "            end_step_pvp = time()\n"
    a3 = "        # Note relevant end of sim values\n"
    a0 = "np.seterr(all='raise')\n"
    for a in (a0, a1, a2, a3):
        assert src.count(a) == 1, a
    src = src.replace(a0, a0 + HOOK)
    src = src.replace(a1, "            _amc_dump('walls', %s)\n" % var + a1)
    src = src.replace(a2, "            _amc_dump('end', %s)\n" % var + a2)
    extra = "completed=completed_paths, completed_x=completed_x_paths, completed_y=completed_y_paths, completed_z=completed_z_paths"
    if kind == "temp":
        extra += ", momentum=momentum_z_change_per_step, e_hot=energy_transfer_hot_per_step, e_cold=energy_transfer_cold_per_step"
    src = src.replace(a3, "        _amc_final(%s)\n" % extra + a3)
    return src


def main():
    kind, k, out = sys.argv[1], int(sys.argv[2]), os.path.abspath(sys.argv[3])
    os.makedirs(out, exist_ok=True)
    work = os.path.join(out, "_run")
    os.makedirs(work, exist_ok=True)
    with open(os.path.join(work, "ref_main.py"), "w") as f:
        f.write(patched_source(kind, k))
    shutil.copy(os.path.join(REF, "utils.py"), os.path.join(work, "utils.py"))
    env = dict(os.environ, PYTHONPATH=os.path.join(HERE, "stubs"), AMC_DUMP_DIR=out)
    with open(os.path.join(out, "stdout.txt"), "w") as log:
        subprocess.check_call([sys.executable, "ref_main.py"], cwd=work, env=env, stdout=log, stderr=subprocess.STDOUT)


if __name__ == "__main__":
    main()
