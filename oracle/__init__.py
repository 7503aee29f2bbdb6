"""CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/ba_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package. The product package camera_calibration_b200 never does.
"""
