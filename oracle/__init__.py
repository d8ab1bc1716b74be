"""CPU oracle -- test infrastructure only (see oracle/model.py and oracle/sim.c headers).

Nothing under elynx_b200/ may import this package."""
