#!/usr/bin/env python
"""Regenerate tests/golden/ by running the UNMODIFIED reference binary (oracle/_ref/papr, built by
oracle/Makefile from /root/reference/papr.c) on every fixture in tests/fixtures.py, in both modes.

    make -C oracle ref && python tests/golden/make_golden.py

Writes  <name>.out / <name>.g.out  (exact stdout) and manifest.json (input md5, rc, stdout md5).
Never hand-edit the outputs.  Needs /root/reference only through the prebuilt oracle/_ref/papr.
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "papr")


def main():
    manifest = {}
    with tempfile.TemporaryDirectory(dir="/dev/shm") as td:
        for name in fixtures.FIXTURES:
            img = fixtures.image(name)
            path = os.path.join(td, name + ".cfile")
            with open(path, "wb") as f:
                f.write(img)
            entry = {"input_md5": fixtures.md5(img), "input_bytes": len(img)}
            for mode, suffix in (("", ".out"), ("-g", ".g.out")):
                cmd = [REF] + ([mode] if mode else []) + [path]
                r = subprocess.run(cmd, capture_output=True, env={"LC_ALL": "C"})
                with open(os.path.join(HERE, name + suffix), "wb") as f:
                    f.write(r.stdout)
                entry["rc" + suffix] = r.returncode
                entry["stdout_md5" + suffix] = fixtures.md5(r.stdout)
                entry["stderr" + suffix] = r.stderr.decode()
            manifest[name] = entry
            print(name, entry["input_bytes"], entry["stdout_md5.out"], entry["stdout_md5.g.out"])
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
