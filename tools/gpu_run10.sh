#!/bin/bash
for v in a b c; do VAR=$v timeout 300 python tools/diag_graph2.py 2>&1 | grep -v Warning | tail -3; done
