#!/bin/bash
# the one-process group step with the NCCL exchange and with peer copies to member 0 (BARTINT_GROUP_EXCHANGE=p2p)
for mode in nccl peer nccl peer; do
  BARTINT_GROUP_EXCHANGE=$mode timeout 200 python tools/group_overhead.py ${1:-2} 1000000 2>&1 | tail -1
done
