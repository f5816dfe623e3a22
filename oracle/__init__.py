"""CPU oracle of the TensoIR hot path — TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and the CPU legs of bench.py; never by tensoir_b200/)."""
