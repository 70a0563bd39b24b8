"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see pbso_oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (openpbso_b200/, include/) never does.
"""
