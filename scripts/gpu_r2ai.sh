#!/bin/bash
ASYNC_OUT=1 timeout 300 python scripts/e2e_gaps.py 2>&1 | tail -1
ASYNC_OUT=0 timeout 300 python scripts/e2e_gaps.py 2>&1 | tail -1
