"""CPU oracle for the CoVA blob-detection hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cova_b200/`` imports this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may use it, and there only as the
checker or as the timed CPU baseline - never as the thing shipped.

It restates, stage by stage, what the reference computes on this path
(all paths relative to the reference tree):

* ``metapreprocess_ref``  cova-rs/gst-plugins/src/metapreprocess/imp.rs:204-236,288-332
* ``blobnet_ref``         utils/model/{preprocessing,encoder,pointwise,decoder,blobnet}.py,
                          utils/train-blobnet.py:57-69,113-119
* ``mask_ref``            config/blobnet/*.txt:26 (segmentation-threshold=0.5) +
                          gst-plugins/gst-maskcopy/gstmaskcopy.cpp:226-230
* ``bboxcc_ref``          cova-rs/gst-plugins/src/bboxcc/process.rs:5-49 (calls the third-party
                          cv::connectedComponentsWithStats, opencv crate 0.53.2, native lib unpinned)
* ``bincode_ref``         cova-rs/bbox/src/bbox.rs:3-29,84-90 (bincode 1.3.3 default options)

Pinning status (see DESIGN.md "Oracle"):
the reference's own tests hold NO golden vectors for any of these stages
(SURVEY.md section 8c).  The CCL restatement is pinned against outputs of
``cv2.connectedComponentsWithStats`` (OpenCV 4.13, the same third-party entry
point ``bboxcc`` calls) committed under ``tests/golden/`` by
``tools/make_golden.py``.  bincode bytes are pinned by hand-computed vectors
from the published bincode 1.3 wire format.  metapreprocess and BlobNet have
no executable reference here (no rustc / GStreamer / TensorFlow): PARITY
UNPINNED for those two restatements.
"""
