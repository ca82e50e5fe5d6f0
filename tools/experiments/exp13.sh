python tools/prof_counters.py
QUERIES=random python tools/prof_counters.py
