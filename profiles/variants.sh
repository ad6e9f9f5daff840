for pad in 0 10000 23000 42000; do
echo "pad $pad"; SMCB_SWEEP_SMEM_PAD=$pad python profiles/fastpath_probe.py
done
