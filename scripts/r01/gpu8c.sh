nvidia-smi topo -m 2>&1 | head -14
lscpu | grep -E "NUMA|Socket|^CPU\(s\)"
for aff in 1 0; do echo "AFF=$aff"; AFF=$aff timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29661 scripts/pcie_probe.py 2>/dev/null | grep "^{"; done
