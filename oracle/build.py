"""ORACLE -- test infrastructure only.  Compiles oracle/tri_oracle.cpp into oracle/libtri_oracle.so.

The reference itself is pure Python whose arithmetic sits in matplotlib's C++ (not vendored,
not installed), so there is nothing under /root/reference that could be compiled into
oracle/_ref/: the reference is "unbuildable" here and the CPU baseline kind is "port".
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "tri_oracle.cpp")
OUT = os.path.join(HERE, "libtri_oracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
