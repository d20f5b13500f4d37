timeout 300 python profiles/trace_cta.py fwd 2>&1 | tail -17
