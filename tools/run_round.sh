timeout 600 python -m pytest tests -m gpu -q -x -k "checkpoint" 2>&1 | tail -12
