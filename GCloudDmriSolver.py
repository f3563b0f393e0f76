#!/usr/bin/env python3
"""Drop-in command line: same flags as the reference's GCloudDmriSolver.py, arithmetic on the B200.
  python GCloudDmriSolver.py -f mesh.npz -M 0 -b 1000 -d 10600 -D 43100 -k 200 -K 3e-3 -gdir 1 0 0"""
import __graft_entry__ as _entry

_entry.load_package()
from dmri_fem_cloud_b200 import cli  # noqa: E402

if __name__ == "__main__":
    cli.main()
