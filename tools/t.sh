python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "auto: $(timeout 300 python tools/quick_run.py config2_phold_1m 1 2>&1 | tail -1)"
for b in 512 480 472 448 416; do
echo "epb $b: $(timeout 300 python tools/quick_run.py config2_phold_1m 1 block_entities=$b 2>&1 | tail -1)"
done
