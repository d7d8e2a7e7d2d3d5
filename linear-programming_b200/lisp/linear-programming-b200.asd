;;;; linear-programming-b200.asd -- the B200 backend for neil-lindquist/linear-programming.
;;;; UNTESTED IN THIS REPO'S CONTAINER: the image has no Common Lisp implementation.  The shim is
;;;; kept deliberately thin; its behaviour is pinned through the Python binding, which goes
;;;; through the identical C ABI (include/b200lp.h).  See INTEGRATION.md.
(defsystem "linear-programming-b200"
  :description "B200 (sm_100a) dense simplex backend for linear-programming's *solver* hook"
  :version "0.1.0"
  :license "MIT"
  :depends-on ("linear-programming" "cffi")
  :components ((:file "b200-solver")))
