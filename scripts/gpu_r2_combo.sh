#!/bin/bash
bash scripts/gpu_r2_hardi_ncu.sh
bash scripts/gpu_r2_l2ahead.sh
