"""Packaging for maggy_b200. The native libraries are built in-tree (``python -m maggy_b200.build`` or
``python setup.py build_native``) with nvcc for sm_100a and loaded through ctypes; nothing is compiled at
``pip install`` time unless ``MAGGY_B200_BUILD_NATIVE=1`` is set."""
import os
import subprocess
import sys

from setuptools import Command, find_packages, setup


class BuildNative(Command):
    description = "compile csrc/ into maggy_b200/lib/*.so (nvcc, sm_100a)"
    user_options = []

    def initialize_options(self):
        pass

    def finalize_options(self):
        pass

    def run(self):
        subprocess.check_call([sys.executable, "-m", "maggy_b200.build"])


if os.environ.get("MAGGY_B200_BUILD_NATIVE") == "1":
    subprocess.check_call([sys.executable, "-m", "maggy_b200.build"])

setup(
    name="maggy_b200",
    version="0.1.0",
    description="B200-native hyperparameter optimisation, ablation and distributed-training framework",
    packages=find_packages(include=["maggy_b200", "maggy_b200.*"]),
    package_data={"maggy_b200": ["lib/*.so", "csrc/**/*"]},
    python_requires=">=3.10",
    install_requires=["torch>=2.6", "numpy", "cloudpickle", "scipy"],
    extras_require={"dev": ["pytest", "pytest-timeout"], "tensorboard": ["tensorboard", "protobuf"]},
    cmdclass={"build_native": BuildNative},
)
