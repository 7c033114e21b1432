"""CPU oracle for the Vision-Kit YOLO detection data path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
CPU baseline -- never as the thing shipped or measured as "ours".  The product
package (``vision_kit_b200``) never imports this package and raises when its
CUDA library is missing.

Layers
------
``restate.py``    pure-numpy restatement of the arithmetic the reference gets
                  from third parties (OpenCV fixed-point bilinear resize,
                  torchvision greedy NMS) plus the reference's own Python logic
                  (letterbox geometry, Detect decode, confidence filter).
``ref_port.py``   the same path written against the same libraries the
                  reference calls (cv2 / torch / torchvision) -- this is what a
                  Vision-Kit user executes on the CPU today and what
                  ``bench.py`` times as the CPU baseline (kind = "port").
``live.py``       imports the *real* reference from ``/root/reference`` (dev
                  container only; the tree does not exist on the GPU box) to pin
                  both of the above and to generate ``tests/golden``.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md §4,
§8c) so the oracle is pinned against outputs of the reference itself, run in
the dev container by ``tests/golden/make_golden.py`` and committed as fixtures.
"""
