#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for s in 1 0; do PTX_STRIP=$s timeout 120 python tools/strip_diag.py 64; PTX_STRIP=$s timeout 120 python tools/strip_diag.py 128; done 2>&1 | tee gpurun_out/r02m_strip_diag.log
