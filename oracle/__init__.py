"""CHECKERS ONLY (test infrastructure): the plain-C restatement of the reference (dfsa_oracle.c),
a runner for the real reference build (oracle/_ref/ref_driver) and an independent dense ground truth.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and only to CHECK or to time the CPU baseline -- the product never routes through it.
"""
