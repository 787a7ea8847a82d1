"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's BSRNN / FlowSE hot path (urgent2026_challenge_track1,
baseline_code/models + sampling + the un-vendored espnet2 pieces).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this package; the product package ``urgent2026_challenge_track1_b200`` never does.
"""
