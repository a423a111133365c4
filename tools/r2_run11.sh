timeout 300 python tools/v2_check.py 99999 2048 2>&1 | grep -E "mixed|fast vs|force evaluation|vs oracle|\[mixed32\]|\[fast\]" | head -16
