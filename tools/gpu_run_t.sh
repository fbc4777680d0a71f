#!/bin/bash
# view store + regression of the pipeline (erosion now runs on smb_view_erode3x3)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 3 gpurun_out/$name.log; grep -E "^(FAILED|ERROR)|^E  |watchdog" gpurun_out/$name.log | head -n 40; }
run t_view tests/test_gpu_view_store.py
run t_pipe tests/test_gpu_pipeline.py -k "not simt"
run t_cli tests/test_gpu_cli.py
