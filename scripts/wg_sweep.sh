for s in 37 74 111 148 296; do
echo "splits $s"; ALIGNN_WG_SPLITS=$s timeout 120 python scripts/probe_wgrad.py 2>&1 | grep "tc=True" 
done
