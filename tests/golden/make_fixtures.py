#!/usr/bin/env python
"""Regenerates the fixtures in this directory from the reference checkout.

The two CSV tables are the reference's own tabulated 1-D solutions that its
shock-tube answer tests compare against (they are test data, not source):

  rj2a_shock_tube_t0.2_res256.csv   input/vlct/MHD_shock_tube/
        Ryu & Jones (1995) fig. 2a MHD shock tube at t = 0.2, 256 cells
        (used by input/vlct/run_MHD_shock_tube_test.py:62-87)
  sod_shock_tube_t0.25_res128.csv   input/vlct/dual_energy_shock_tube/
        Sod shock tube at t = 0.25, 128 cells
        (used by input/vlct/run_dual_energy_shock_tube_test.py:46-78)

usage: python tests/golden/make_fixtures.py [/root/reference]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["input/vlct/MHD_shock_tube/rj2a_shock_tube_t0.2_res256.csv",
         "input/vlct/dual_energy_shock_tube/sod_shock_tube_t0.25_res128.csv"]


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    for rel in FILES:
        shutil.copyfile(os.path.join(ref, rel),
                        os.path.join(HERE, os.path.basename(rel)))
        print("copied", rel)


if __name__ == "__main__":
    main()
