"""TEST INFRASTRUCTURE ONLY.

``oracle/`` holds a CPU restatement of the reference Hex simulator
(MBPrdctns/hex_gym_env, ``minihex/{HexGame,HexSingleGame,SelfplayWrapper,__init__}.py``)
used to check the CUDA path. Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it. The
product package ``hex_gym_env_b200`` never imports anything from here.
"""
