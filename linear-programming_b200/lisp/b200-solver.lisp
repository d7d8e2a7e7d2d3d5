;;;; b200-solver.lisp -- CFFI shim: (setf linear-programming:*solver* 'linear-programming-b200:b200-solver)
;;;;
;;;; Contract served: src/solver.lisp:39-56 of the reference -- a function (problem &key ...) that
;;;; returns an object answering solution-problem / solution-objective-value / solution-variable /
;;;; solution-reduced-cost.  The object returned IS the reference's own `tableau` (built by its
;;;; exported build-tableau, src/simplex.lisp:142), with the solved right-hand-side column,
;;;; objective row and basis written into it, so the existing methods (src/solver.lisp:61-80) and
;;;; both with-* macros work unchanged.  Only n-solve-tableau (src/simplex.lisp:399-461) is
;;;; replaced: it runs in libb200lp.so on the GPU.
;;;;
;;;; UNTESTED HERE (no Lisp in the build image); mirrors linear-programming_b200/simplex.py, which
;;;; is tested through the same C ABI.
(defpackage :linear-programming-b200
  (:use :cl)
  (:import-from :linear-programming/problem
                #:problem-type #:problem-vars #:problem-objective-var #:problem-objective-func
                #:problem-integer-vars #:problem-var-bounds #:problem-constraints)
  (:import-from :linear-programming/simplex
                #:build-tableau #:tableau-matrix #:tableau-basis-columns #:tableau-var-count
                #:tableau-constraint-count #:tableau-instance-problem #:tableau-problem
                #:tableau-variable #:tableau-objective-value)
  (:import-from :linear-programming/conditions
                #:solver-error #:unbounded-problem-error #:infeasible-problem-error)
  (:export #:b200-solver #:install #:b200-error))
(in-package :linear-programming-b200)

(cffi:define-foreign-library libb200lp
  (t (:default "libb200lp")))          ; put linear-programming_b200/ on the loader path
(cffi:use-foreign-library libb200lp)

;;; include/b200lp.h ---------------------------------------------------------------------------
(cffi:defcstruct opts
  (fp-tolerance-factor :double) (pivot-rule :int32) (writeback-full :int32) (max-iters :int64)
  (ndev :int32) (devices :int32 :count 8) (trace-capacity :int32) (poll-interval :int32)
  (time-kernels :int32) (pivot-variant :int32) (reserved :int32 :count 6))

(cffi:defcstruct result
  (status :int32) (n-devices :int32) (iterations :int64) (iterations-phase1 :int64)
  (iterations-cleanup :int64) (objective :double) (ms-total :double) (ms-h2d :double)
  (ms-solve :double) (ms-d2h :double) (ms-pivot-kernel :double) (pivot-kernel-launches :int64)
  (kernel-launches :int64) (h2d-bytes :int64) (d2h-bytes :int64) (bytes-per-pivot :int64)
  (trace-len :int32) (exchange-mode :int32) (ms-look-kernel :double) (ms-exchange :double)
  (look-kernel-launches :int64))

(cffi:defcfun ("b200lp_solve" %solve) :int
  (opts :pointer) (tab :pointer) (r :int64) (c :int64) (ld :int64) (basis :pointer)
  (is-max :int32) (out :pointer) (trace-j :pointer) (trace-r :pointer))

(cffi:defcfun ("b200lp_solve_two_phase" %solve-two-phase) :int
  (opts :pointer) (art-tab :pointer) (c-art :int64) (ld-art :int64) (art-basis :pointer)
  (main-tab :pointer) (r :int64) (c :int64) (ld :int64) (main-basis :pointer)
  (is-max :int32) (out :pointer))

(cffi:defcfun ("b200lp_strerror" %strerror) :string (code :int))
(cffi:defcfun ("b200lp_last_error" %last-error) :string)

(define-condition b200-error (solver-error)
  ((code :initarg :code :reader b200-error-code))
  (:report (lambda (c s) (format s "libb200lp: ~A (~D) ~A" (%strerror (b200-error-code c))
                                 (b200-error-code c) (%last-error)))))

(defun signal-status (status)
  "include/b200lp.h status codes -> src/conditions.lisp:43-77"
  (case status
    (0 nil)
    (1 (error 'unbounded-problem-error))
    (2 (error 'infeasible-problem-error))
    (t (error 'b200-error :code status))))

;;; tableau <-> flat fp64 buffers ---------------------------------------------------------------
(defun export-tableau (tableau)
  "Row-major double-float copy of the boxed (simple-array real 2) matrix plus an int32 basis."
  (let* ((matrix (tableau-matrix tableau))
         (rows (array-dimension matrix 0)) (cols (array-dimension matrix 1))
         (tab (cffi:foreign-alloc :double :count (* rows cols)))
         (basis (cffi:foreign-alloc :int32 :count (max 1 (1- rows)))))
    (dotimes (k (* rows cols))
      (setf (cffi:mem-aref tab :double k) (coerce (row-major-aref matrix k) 'double-float)))
    (dotimes (i (1- rows))
      (setf (cffi:mem-aref basis :int32 i) (aref (tableau-basis-columns tableau) i)))
    (values tab basis rows cols)))

(defun import-solution (tableau tab basis rows cols)
  "What the accessors read (src/simplex.lisp:74-120): RHS column, objective row, basis."
  (let ((matrix (tableau-matrix tableau)))
    (dotimes (i rows)
      (setf (aref matrix i (1- cols)) (cffi:mem-aref tab :double (+ (* i cols) (1- cols)))))
    (dotimes (j cols)
      (setf (aref matrix (1- rows) j) (cffi:mem-aref tab :double (+ (* (1- rows) cols) j))))
    (dotimes (i (1- rows))
      (setf (aref (tableau-basis-columns tableau) i) (cffi:mem-aref basis :int32 i)))
    tableau))

(defun fill-opts (o tolerance devices pivot-rule max-iterations)
  "Zero the b200lp_opts at O (every zero field = library default), then set what was asked for."
  (dotimes (k (cffi:foreign-type-size '(:struct opts)))
    (setf (cffi:mem-aref o :uint8 k) 0))
  (flet ((slot (name value)
           (setf (cffi:foreign-slot-value o '(:struct opts) name) value)))
    (slot 'fp-tolerance-factor (coerce tolerance 'double-float))
    (slot 'pivot-rule (or pivot-rule 0))
    (slot 'max-iters (or max-iterations 0))
    (slot 'ndev (if (> (length devices) 1) (length devices) 0)))
  (let ((dev (cffi:foreign-slot-pointer o '(:struct opts) 'devices)))
    (loop for d in devices
          for k from 0 below 8
          do (setf (cffi:mem-aref dev :int32 k) d))))

(defun gpu-solve-tableau (tableau &key (tolerance 1024) devices pivot-rule max-iterations)
  "n-solve-tableau (src/simplex.lisp:399-461) on the GPU.  TABLEAU is a tableau or the
(art main) list build-tableau returns; the (main) tableau is returned, solved."
  (let* ((two-phase (listp tableau))
         (main (if two-phase (second tableau) tableau))
         (is-max (if (eq 'max (problem-type (tableau-instance-problem main))) 1 0)))
    (cffi:with-foreign-objects ((o '(:struct opts)) (res '(:struct result)))
      (fill-opts o tolerance devices pivot-rule max-iterations)
      (multiple-value-bind (tab basis rows cols) (export-tableau main)
        (unwind-protect
             (progn
               (if two-phase
                   (multiple-value-bind (atab abasis arows acols) (export-tableau (first tableau))
                     (declare (ignore arows))
                     (unwind-protect
                          (signal-status (%solve-two-phase o atab acols acols abasis
                                                           tab rows cols cols basis is-max res))
                       (cffi:foreign-free atab)
                       (cffi:foreign-free abasis)))
                   (signal-status (%solve o tab rows cols cols basis is-max res
                                          (cffi:null-pointer) (cffi:null-pointer))))
               (import-solution main tab basis rows cols))
          (cffi:foreign-free tab)
          (cffi:foreign-free basis))))))

;;; branch and bound: control flow of simplex-solver (src/simplex.lisp:506-542) -------------------
(defun integral-p (value tolerance)
  "(integerp value) for exact rationals (src/simplex.lisp:479); on double-floats the same test
with the library's own fp= tolerance (src/utils.lisp:84-93)."
  (<= (abs (- value (fround value))) (* tolerance double-float-epsilon)))

(defun violated-integer-constraint (tableau tolerance)
  (dolist (var (problem-integer-vars (tableau-problem tableau)))
    (unless (integral-p (tableau-variable tableau var) tolerance)
      (return var))))

(defun build-and-solve (problem extra tolerance backend-args)
  (handler-case
      (apply #'gpu-solve-tableau
             (build-tableau problem
                            (if (null extra)
                                problem
                                (linear-programming/problem::make-problem
                                 :type (problem-type problem)
                                 :vars (problem-vars problem)
                                 :objective-var (problem-objective-var problem)
                                 :objective-func (problem-objective-func problem)
                                 :integer-vars (problem-integer-vars problem)
                                 :var-bounds (problem-var-bounds problem)
                                 :constraints (append extra (problem-constraints problem))))
                            :fp-tolerance-factor tolerance)
             backend-args)
    (infeasible-problem-error () :infeasible)))

(defun b200-solver (problem &key (fp-tolerance 1024) devices pivot-rule max-iterations
                    &allow-other-keys)
  "The *solver* backend function.  Keywords: :fp-tolerance (as simplex-solver, src/simplex.lisp:511),
:devices (list of CUDA ordinals; more than one row-block shards the tableau), :pivot-rule
(0 reference rule, 1 Bland), :max-iterations."
  (let ((backend-args (list :tolerance fp-tolerance :devices devices :pivot-rule pivot-rule
                            :max-iterations max-iterations))
        (better (if (eq (problem-type problem) 'max) #'< #'>))
        (best nil) (solution nil) (stack (list '())))
    (loop while stack
          do (let* ((entry (pop stack))
                    (tab (build-and-solve problem entry fp-tolerance backend-args)))
               (unless (eq tab :infeasible)
                 (let ((violated (violated-integer-constraint tab fp-tolerance))
                       (value (tableau-objective-value tab)))
                   (cond
                     ((and violated best (not (funcall better best value))))
                     (violated
                      (let ((val (tableau-variable tab violated)))
                        (setf stack (list* (list* `(<= ((,violated . 1)) ,(floor val)) entry)
                                           (list* `(>= ((,violated . 1)) ,(ceiling val)) entry)
                                           stack))))
                     ((or (null best) (funcall better best value))
                      (setf best value solution tab)))))))
    (or solution (error 'infeasible-problem-error))))

(defun install ()
  "Make the B200 backend the default for solve-problem / with-solved-problem."
  (setf linear-programming/solver:*solver* 'b200-solver))
