"""btfem: B200-native Bloch-Torrey FEM time-stepper (the hot path of van-dang/DMRI-FEM-Cloud).

csrc/        CUDA kernels (sm_100a) + the C-ABI of include/btfem.h  -> libbtfem.so
btfem.py     ctypes binding
dmrifemlib.py  mirror of the reference's DmriFemLib operator interface on top of the binding
meshes.py    mesh readers (gmsh v2, DOLFIN XML) and synthetic generators
cli.py       GCloudDmriSolver.py-compatible command line

The directory name contains hyphens, so import it through __graft_entry__.load_package()
(registers it as `dmri_fem_cloud_b200`)."""
