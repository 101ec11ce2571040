for d in 64 128 192; do for w in 128 192; do
echo "BN_D=$d BN_WIDE=$w"; MDTB200_BN_D=$d MDTB200_BN_WIDE=$w python tools/train_once.py 2>&1 | grep -o '"eager": {"value": [0-9.]*, "ms_per_step": [0-9.]*\|"graphed": {"value": [0-9.]*, "ms_per_step": [0-9.]*\|loss_last": [0-9.]*'
done; done
