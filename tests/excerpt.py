"""Loaders over tests/golden/intel_excerpt.txt (the first 45 sweeps of the
reference's data/intel.txt with their ODOM lines)."""
import os

from thesis_b200 import loaders

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class ExcerptLidar(loaders.IntelLidarData):
    FILE = "intel_excerpt.txt"

    def __init__(self):
        super().__init__(GOLDEN)


class ExcerptIMU(loaders.IntelIMUData):
    FILE = "intel_excerpt.txt"

    def __init__(self):
        super().__init__(GOLDEN)
