python bench.py --skip-cpu --text-index --steps 2 --warmup 3 --sweep 9,9 "$@" 2>&1 >/dev/null | grep -E "sweep|unique-walk"
