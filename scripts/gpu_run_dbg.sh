#!/bin/bash
for v in 1 2 7 3 4 5 6 8 9; do timeout 60 ./scripts/tma_probe $v 2>&1 | tail -2; done
