import sys
sys.path.insert(0, '.')
from orbit_b200 import lib as L
import pytest
k, v = sys.argv[1].split('=')
assert L.load().orbit_set_global_option(k.encode(), int(v)) == 0
sys.exit(pytest.main(sys.argv[2:]))
