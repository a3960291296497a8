"""CPU oracle for the InstantAvatar hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this package.  Nothing under instantavatar_b200/ imports it.
"""
