"""CPU oracle for the RAM-Net hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``rpg_ramnet_b200/`` may import this
package.  The only legitimate importers are ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs, and there only as the checker / the CPU baseline, never as
the thing shipped.

Parity pinning: the reference repository holds NO golden vectors, known-answer
tests or fixtures for this path.  The oracle is therefore
pinned against outputs of the reference itself, executed in the build
container by ``oracle/make_golden.py`` (which imports the reference Python
modules from ``/root/reference/RAM_Net``) and committed as small fixtures
under ``tests/golden/``.
"""
