"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the MT-DREAM(ZS) step path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import, link or execute it, and only as the checker / the CPU arm.
The product package ``pydream_b200`` never imports this package.
"""
