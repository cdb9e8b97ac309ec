"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see als_oracle.c header).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this.
"""
