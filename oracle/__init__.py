"""TEST INFRASTRUCTURE ONLY.  CPU oracles for the solve path; imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package."""
