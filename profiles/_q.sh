python profiles/nms_phases.py 2>&1 | grep -v Warn | grep -v "   page" | head -16
python profiles/time_inference.py 2>&1 | grep -v Warn | tail -2
