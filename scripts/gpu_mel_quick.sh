#!/bin/bash
# quick GPU session (tests + timing + bench line, no ncu)
bash scripts/gpu_r02.sh
