#!/usr/bin/env python
"""Probe for the reference's real solver stack (CasADi -> IPOPT -> MUMPS) on the machine this runs on.

SURVEY.md 8c plan step 1 / BASELINE.md 2 step 1: before falling back to the restated oracle, actually try
`import casadi` on the GPU box, look for a driver-side install under baseline/_ref, look for wheels in the
offline wheelhouse, and record what pip says when asked for the package.  The output (one JSON document) is
committed under profiles/ so the "CasADi is not available" statement in DESIGN.md rests on a record.

    python tools/probe_casadi.py > gpurun_out/probe_casadi.json
"""
import glob
import importlib
import json
import os
import platform
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def try_import(name, extra_path=None):
    if extra_path and extra_path not in sys.path:
        sys.path.insert(0, extra_path)
    try:
        m = importlib.import_module(name)
        return {"ok": True, "version": getattr(m, "__version__", None), "file": getattr(m, "__file__", None)}
    except Exception as e:      # ModuleNotFoundError, ImportError of a missing shared object, ...
        return {"ok": False, "error": "%s: %s" % (type(e).__name__, e)}


def run(cmd, timeout=60):
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout, text=True)
        return {"cmd": " ".join(cmd), "rc": p.returncode, "tail": p.stdout.strip().splitlines()[-6:]}
    except Exception as e:
        return {"cmd": " ".join(cmd), "rc": None, "tail": ["%s: %s" % (type(e).__name__, e)]}


def main():
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    rec = {
        "host": platform.node(), "python": sys.version.split()[0], "cpu_count": os.cpu_count(),
        "import_casadi": try_import("casadi"),
        "import_casadi_from_baseline_ref": try_import("casadi", ref_dir) if os.path.isdir(ref_dir) else {"ok": False, "error": "baseline/_ref does not exist"},
        "import_cyipopt": try_import("cyipopt"),
        "import_ipopt": try_import("ipopt"),
        "import_rospy": try_import("rospy"),
        "wheelhouse": sorted(os.path.basename(p) for pat in ("*casadi*", "*ipopt*", "*mumps*", "*coinhsl*")
                             for p in glob.glob(os.path.join("/opt/wheelhouse", pat))),
        "wheelhouse_count": len(glob.glob("/opt/wheelhouse/*")),
        "shared_objects": sorted(p for pat in ("libipopt*", "libcasadi*", "libdmumps*", "libcoinmumps*")
                                 for d in ("/usr/lib", "/usr/lib/x86_64-linux-gnu", "/usr/local/lib", "/opt")
                                 for p in glob.glob(os.path.join(d, "**", pat), recursive=True))[:20],
        "which_ipopt": run(["sh", "-c", "command -v ipopt || echo 'ipopt: not found'"]),
        "pip_download": run([sys.executable, "-m", "pip", "download", "--no-deps", "-d", "/tmp/_probe_casadi", "casadi"], timeout=120),
        "pip_install_from_wheelhouse": run([sys.executable, "-m", "pip", "install", "--no-index", "--find-links", "/opt/wheelhouse",
                                            "--target", "/tmp/_probe_casadi_t", "casadi"], timeout=120),
    }
    rec["casadi_available"] = bool(rec["import_casadi"]["ok"] or rec["import_casadi_from_baseline_ref"]["ok"])
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
