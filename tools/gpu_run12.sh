#!/bin/bash
for m in global thread_local relaxed; do MODE=$m timeout 300 python tools/diag_graph3.py 2>&1 | grep -v "Warning\|detach" | tail -9; done
WU=0 MODE=global timeout 300 python tools/diag_graph3.py 2>&1 | grep -v "Warning\|detach" | tail -4
