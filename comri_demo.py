#!/usr/bin/env python3
"""The comri FEniCS-HPC demos (comri/{one-comp,two-comp,multilayer}/hpc-fenics-cpp/main.cpp) on the B200: same flags.
  python comri_demo.py one-comp -m cyl12_r_3E_6_vol.msh.zip -b 1000 -d 10600 -D 43100 -k 200 -v 1 0 0 -K 3e-3
  python comri_demo.py two-comp -m multi_layer_torus.xml.zip -c multi_layer_torus_compt1.xml.zip -b 4000 -p 5e-5"""
import sys

import __graft_entry__ as _entry

_entry.load_package()
from dmri_fem_cloud_b200 import comri  # noqa: E402

if __name__ == "__main__":
    sys.exit(comri.main())
