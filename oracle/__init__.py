"""oracle/ -- CPU restatement of the reference's hot path and the compiled reference itself.
TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by eigencuda_b200/."""
