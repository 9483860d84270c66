timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | cut -c1-300
