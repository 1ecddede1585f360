for nc in 2 3 4; do for ramp in 1 0; do
echo "NC=$nc RAMP=$ramp"; CB2_NC=$nc CB2_RAMP=$ramp python tools/e2e_probe.py 262144 524288 2>&1 | tail -2
done; done
