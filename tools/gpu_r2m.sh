#!/bin/bash
set -u
timeout 600 python tools/bench_rank_share.py 2>&1 | tail -n 4
