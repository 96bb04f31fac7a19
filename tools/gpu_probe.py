"""Runs groups of GPU parity cases in separate subprocesses (a CUDA fault in one group must not hide
the others) and writes one JSON line per case to gpurun_out/probe.jsonl.

    python tools/gpu_probe.py                 # all groups
    python tools/gpu_probe.py --group wgrad   # one group, in-process
"""
import argparse
import json
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")


def run_group(group):
    import torch
    import conv_cases as cc
    from cycle_depth_estimation_b200 import _lib
    table = {"fwd": (cc.FWD_CASES, cc.conv_fwd_case), "rowpack": (cc.ROWPACK_CASES, cc.conv_rowpack_case),
             "wgrad": (cc.WGRAD_CASES, cc.conv_wgrad_case),
             "wgrad2": (cc.WGRAD_FEWCOUT_CASES, cc.conv_wgrad_fewcout_case)}
    cases, fn = table[group]
    for name, kw in cases.items():
        rec = {"group": group, "case": name}
        t0 = time.time()
        try:
            rec.update(fn(**kw))
        except Exception as e:  # noqa: BLE001
            rec.update({"ok": False, "exc": "%s: %s" % (type(e).__name__, e), "tb": traceback.format_exc()[-200:]})
        try:
            rec["abort_flag"] = _lib.lib().cdb_device_abort_flag()
        except Exception as e:  # noqa: BLE001
            rec["abort_flag"] = str(e)
        rec["sec"] = round(time.time() - t0, 3)
        print(json.dumps(rec), flush=True)
        sys.stderr.write("%-34s %s err=%s %s\n" % (name, "OK  " if rec.get("ok") else "FAIL", rec.get("err"), str(rec.get("exc", ""))[:160]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default=None)
    ap.add_argument("--groups", default="fwd,rowpack,wgrad,wgrad2")
    args = ap.parse_args()
    if args.group:
        run_group(args.group)
        return
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "probe.jsonl"), "a") as log:
        for group in args.groups.split(","):
            try:
                proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", group],
                                      capture_output=True, text=True, timeout=600)
                out, err, rc = proc.stdout, proc.stderr, proc.returncode
            except subprocess.TimeoutExpired as e:
                out, err, rc = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "TIMEOUT", -9
            log.write(out)
            log.write(json.dumps({"group": group, "rc": rc, "stderr_tail": err[-1500:]}) + "\n")
            log.flush()
            print("== group %s rc=%s\n%s" % (group, rc, err[-4000:]))


if __name__ == "__main__":
    main()
