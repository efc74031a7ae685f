python profiles/nms_phases.py 2>&1 | grep -v Warn | head -90
