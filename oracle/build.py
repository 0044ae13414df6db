"""Compile the C part of the oracle (TEST INFRASTRUCTURE - see oracle/__init__.py).

The reference (Iselix/kwiiyatta) has no C/C++ sources of its own, so there is no
``oracle/_ref`` build: the only native oracle code is this repo's own restatement.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(BUILD_DIR, 'liboracle_dtw.so')
SOURCES = [os.path.join(HERE, 'dtw_c.c')]


def build(force=False):
    os.makedirs(BUILD_DIR, exist_ok=True)
    if not force and os.path.exists(LIB):
        newest = max(os.path.getmtime(s) for s in SOURCES)
        if os.path.getmtime(LIB) >= newest:
            return LIB
    cmd = ['gcc', '-O2', '-ffp-contract=off', '-mfma', '-shared', '-fPIC',
           '-o', LIB] + SOURCES + ['-lm']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force=True))
