"""CPU oracle for the fibergen Lippmann-Schwinger solve loop.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the timed CPU
baseline.  The shipped path (``fibergen_b200``) never imports this package and has
no CPU fallback.

Parity pinning status (see DESIGN.md "Oracle"): the reference (one 27k-line C++
translation unit needing Boost, FFTW3, LAPACK bindings, libpng) cannot be built in
this image and ships no golden vectors.  The oracle is therefore pinned against the
reference's own *known answers*: the operator identities of ``fibergen --test``
(fg:23946-23974, fg:24086-24182, fg:24460-24583), the closed-form laminate of
``demo/elasticity/laminate`` (fg:26405-26474), the Hashin coated-sphere value
documented in ``demo/elasticity/hashin/project.xml:30-32`` and the homogeneous
one-iteration case.  Absolute DFT values are not pinned by any reference test
(FFTW is an un-vendored dependency); numpy/scipy pocketfft is used as the FP64 DFT.
"""
