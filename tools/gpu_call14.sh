#!/bin/bash
out=gpurun_out; mkdir -p $out
export SCNERF_LIB=$PWD/scnerf_b200/csrc/libscnerf_b200_timeline.so
timeout 300 python tools/timeline_pipe.py bf16x3 > $out/r2n_fwd_pipe_timeline_x3_infer.txt 2>&1
timeout 300 python tools/timeline_pipe.py bf16x3 train > $out/r2n_fwd_pipe_timeline_x3_train.txt 2>&1
grep -v Warn $out/r2n_fwd_pipe_timeline_x3_infer.txt | head -12 | cut -c100-400; grep -v Warn $out/r2n_fwd_pipe_timeline_x3_train.txt | head -12 | cut -c100-400
