set -x
timeout 900 python -m pytest tests/test_host_classes.py tests/test_mesh_reference.py -m gpu -x -q -s 2>&1 | tail -15
