for v in "" ring "" ring; do
  if [ -n "$v" ]; then export RN_B200_LIB=$PWD/retinanet-for-table-detection_b200/librn_b200.$v.so; else unset RN_B200_LIB; fi
  echo "== variant '$v'"
  python bench.py --config 2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('k3', d['roofline']['us_per_launch'], d['roofline']['frac'], 'nms', d['nms']['us_per_launch'], 'inference', d['inference']['reference_semantics']['pages_per_s'])"
done
export RN_B200_LIB=$PWD/retinanet-for-table-detection_b200/librn_b200.ring.so
python -m pytest tests -m gpu -x -q -k "filter or detect or config or nms" 2>&1 | tail -2
