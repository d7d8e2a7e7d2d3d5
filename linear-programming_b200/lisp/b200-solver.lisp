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
;;;; Two ways in:
;;;;   * default: BUILD-TABLEAU-F64 (below) fills foreign double-float buffers straight from the
;;;;     `problem` struct -- the boxed (simple-array real 2) of src/simplex.lisp:214-221 is never
;;;;     built (at m=8192, n=16384 that is 201 M boxed cells before the GPU sees a byte) -- and the
;;;;     solve returns a B200-SOLUTION answering the four solution-* generics;
;;;;   * :boxed t: the reference's own build-tableau + a cell-by-cell coerce; returns the
;;;;     reference's `tableau` object (kept for the no-constraint path and as a cross-check).
;;;;
;;;; UNTESTED HERE (no Lisp in the build image); mirrors linear-programming_b200/simplex.py, which
;;;; is tested through the same C ABI.  The struct layouts below are checked against the library
;;;; at load time (b200lp_abi_sizes).
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
                #:solver-error #:unbounded-problem-error #:infeasible-problem-error #:parsing-error)
  (:import-from :linear-programming/solver
                #:solution-problem #:solution-objective-value #:solution-variable
                #:solution-reduced-cost)
  (:export #:b200-solver #:install #:b200-error #:b200-solution #:build-tableau-f64))
(in-package :linear-programming-b200)

(cffi:define-foreign-library libb200lp
  (t (:default "libb200lp")))          ; put linear-programming_b200/ on the loader path
(cffi:use-foreign-library libb200lp)

;;; include/b200lp.h ---------------------------------------------------------------------------
(cffi:defcstruct opts
  (fp-tolerance-factor :double) (pivot-rule :int32) (writeback-full :int32) (max-iters :int64)
  (ndev :int32) (devices :int32 :count 8) (trace-capacity :int32) (poll-interval :int32)
  (time-kernels :int32) (pivot-variant :int32) (feas-mode :int32) (reserved :int32 :count 5))

(cffi:defcstruct result
  (status :int32) (n-devices :int32) (iterations :int64) (iterations-phase1 :int64)
  (iterations-cleanup :int64) (objective :double) (ms-total :double) (ms-h2d :double)
  (ms-solve :double) (ms-d2h :double) (ms-pivot-kernel :double) (pivot-kernel-launches :int64)
  (kernel-launches :int64) (h2d-bytes :int64) (d2h-bytes :int64) (bytes-per-pivot :int64)
  (trace-len :int32) (exchange-mode :int32) (ms-look-kernel :double) (ms-exchange :double)
  (look-kernel-launches :int64) (loop-mode :int32) (look-ctas :int32) (ms-look-wait :double)
  (ms-look-ratio :double) (ms-look-push :double) (ms-look-peer-wait :double) (ms-look-row :double)
  (sm-clock-mhz :double) (ms-look-dbg :double :count 8) (redundant-rows :int64)
  (look-cluster :int32) (reserved-r :int32))

(cffi:defcfun ("b200lp_solve" %solve) :int
  (opts :pointer) (tab :pointer) (r :int64) (c :int64) (ld :int64) (basis :pointer)
  (is-max :int32) (out :pointer) (trace-j :pointer) (trace-r :pointer))

(cffi:defcfun ("b200lp_solve_two_phase" %solve-two-phase) :int
  (opts :pointer) (art-tab :pointer) (c-art :int64) (ld-art :int64) (art-basis :pointer)
  (main-tab :pointer) (r :int64) (c :int64) (ld :int64) (main-basis :pointer)
  (is-max :int32) (out :pointer))

(cffi:defcfun ("b200lp_abi_sizes" %abi-sizes) :void (opts-size :pointer) (result-size :pointer))
(cffi:defcfun ("b200lp_strerror" %strerror) :string (code :int))
(cffi:defcfun ("b200lp_last_error" %last-error) :string)

(define-condition b200-error (solver-error)
  ((code :initarg :code :reader b200-error-code))
  (:report (lambda (c s) (format s "libb200lp: ~A (~D) ~A" (%strerror (b200-error-code c))
                                 (b200-error-code c) (%last-error)))))

;; the hand-written layouts above must be the library's (include/b200lp.h)
(cffi:with-foreign-objects ((so :int64) (sr :int64))
  (%abi-sizes so sr)
  (assert (= (cffi:mem-ref so :int64) (cffi:foreign-type-size '(:struct opts))) ()
          "b200lp_opts is ~D bytes in libb200lp, ~D here" (cffi:mem-ref so :int64)
          (cffi:foreign-type-size '(:struct opts)))
  (assert (= (cffi:mem-ref sr :int64) (cffi:foreign-type-size '(:struct result))) ()
          "b200lp_result is ~D bytes in libb200lp, ~D here" (cffi:mem-ref sr :int64)
          (cffi:foreign-type-size '(:struct result))))

(defun signal-status (status)
  "include/b200lp.h status codes -> src/conditions.lisp:43-77"
  (case status
    (0 nil)
    (1 (error 'unbounded-problem-error))
    (2 (error 'infeasible-problem-error))
    ;; 3 iteration limit, 4 / 5 the reference's two plain `error`s about a basic artificial
    ;; (src/simplex.lisp:423-424, 432-433), negative = CUDA / NCCL / argument failures
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

(defun fill-opts (o tolerance devices pivot-rule max-iterations &optional feas-mode)
  "Zero the b200lp_opts at O (every zero field = library default), then set what was asked for."
  (dotimes (k (cffi:foreign-type-size '(:struct opts)))
    (setf (cffi:mem-aref o :uint8 k) 0))
  (flet ((slot (name value)
           (setf (cffi:foreign-slot-value o '(:struct opts) name) value)))
    (slot 'fp-tolerance-factor (coerce tolerance 'double-float))
    (slot 'pivot-rule (or pivot-rule 0))
    (slot 'max-iters (or max-iterations 0))
    (slot 'feas-mode (or feas-mode 0))           ; 0 scaled (default), 1 the reference's literal tests
    (slot 'ndev (if (> (length devices) 1) (length devices) 0)))
  (let ((dev (cffi:foreign-slot-pointer o '(:struct opts) 'devices)))
    (loop for d in devices
          for k from 0 below 8
          do (setf (cffi:mem-aref dev :int32 k) d))))

(defun gpu-solve-tableau (tableau &key (tolerance 1024) devices pivot-rule max-iterations feas-mode)
  "n-solve-tableau (src/simplex.lisp:399-461) on the GPU.  TABLEAU is a tableau or the
(art main) list build-tableau returns; the (main) tableau is returned, solved."
  (let* ((two-phase (listp tableau))
         (main (if two-phase (second tableau) tableau))
         (is-max (if (eq 'max (problem-type (tableau-instance-problem main))) 1 0)))
    (cffi:with-foreign-objects ((o '(:struct opts)) (res '(:struct result)))
      (fill-opts o tolerance devices pivot-rule max-iterations feas-mode)
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

;;; direct fp64 build (SURVEY 8 f2): src/simplex.lisp:189-325 without the boxed matrix -----------
(defstruct f64-tableau
  tab basis (rows 0 :type fixnum) (cols 0 :type fixnum))     ; foreign :double / :int32 buffers

(defun alloc-f64 (rows cols)
  (make-f64-tableau
   :tab (cffi:foreign-alloc :double :count (* rows cols) :initial-element 0d0)
   :basis (cffi:foreign-alloc :int32 :count (max 1 (1- rows)) :initial-element 0)
   :rows rows :cols cols))

(defun free-f64 (ft)
  (when ft
    (cffi:foreign-free (f64-tableau-tab ft))
    (cffi:foreign-free (f64-tableau-basis ft))))

(declaim (inline f64-set))
(defun f64-set (ft row col value)
  "VALUE is computed in the input's own numeric type (exact for integers and ratios, as in the
reference) and rounded once, on the way into the buffer."
  (setf (cffi:mem-aref (f64-tableau-tab ft) :double (+ (* row (f64-tableau-cols ft)) col))
        (coerce value 'double-float)))

(defun build-tableau-f64 (problem instance-problem)
  "build-tableau (src/simplex.lisp:142-328) writing double-floats straight into foreign memory.
Returns (values main art mappings): MAIN and ART are F64-TABLEAUs (ART is nil when the slack basis
is feasible); MAPPINGS maps a variable to (:positive col offset) | (:negative col offset) |
(:signed col), as src/simplex.lisp:189-212.  Returns NIL for a problem without constraints
(:153-186) -- the caller takes the boxed path for that one."
  (let* ((constraints (copy-list (problem-constraints instance-problem)))
         (vars (problem-vars problem))
         (mappings (make-hash-table :test 'eq :size (max 16 (* 2 (length vars)))))
         (num-var-cols 0))
    (when (null constraints)
      (return-from build-tableau-f64 nil))
    ;; variable -> column(s), bounded variables add a row  :189-212
    (let ((column 0))
      (loop for var across vars
            for bound = (find var (problem-var-bounds problem) :key #'first)
            do (setf (gethash var mappings)
                     (cond
                       ((null bound) (list :positive column 0))
                       ((and (cadr bound) (cddr bound))
                        (push (if (<= 0 (cddr bound))
                                  `(<= ((,var . 1)) ,(cddr bound))
                                  `(>= ((,var . 1)) ,(- (cddr bound))))   ; kept as the reference writes it
                              constraints)
                        (list :positive column (cadr bound)))
                       ((cadr bound) (list :positive column (cadr bound)))
                       ((cddr bound) (list :negative column (cddr bound)))
                       (t (prog1 (list :signed column) (incf column)))))
               (incf column))
      (setf num-var-cols column))
    (let* ((m (length constraints))
           (num-slack (count-if-not (lambda (c) (eq (first c) '=)) constraints))
           (num-cols (+ num-var-cols num-slack 1))
           (main (alloc-f64 (1+ m) num-cols))
           (row-cells (make-array m :initial-element nil)) ; exact (col . value) per row, RHS included
           (art-rows '())                                  ; pushed: reverse row order, as :258
           (col-offset 0)
           (art nil)
           (ok nil))
      (unwind-protect
           (flet ((set-basis (ft row col)
                    (setf (cffi:mem-aref (f64-tableau-basis ft) :int32 row) col)))
             ;; constraint rows  :223-265
             (loop for row from 0 below m
                   for constraint in constraints
                   do (let ((op (first constraint)) (rhs (third constraint)) (cells '()))
                        (loop for (var . coef) in (second constraint)
                              for mp = (gethash var mappings)
                              do (ecase (first mp)
                                   (:positive (push (cons (second mp) coef) cells)
                                    (decf rhs (* coef (third mp))))
                                   (:negative (push (cons (second mp) (- coef)) cells)
                                    (decf rhs (* coef (third mp))))
                                   (:signed (push (cons (second mp) coef) cells)
                                    (push (cons (1+ (second mp)) (- coef)) cells))))
                        (when (< rhs 0)                     ; keep the RHS non-negative  :243-252
                          (setf cells (mapcar (lambda (c) (cons (car c) (- (cdr c)))) cells)
                                rhs (- rhs)
                                op (case op (<= '>=) (>= '<=) (t op))))
                        (case op
                          (<= (push (cons (+ num-var-cols col-offset) 1) cells)
                           (set-basis main row (+ num-var-cols col-offset))
                           (incf col-offset))
                          (>= (push row art-rows)
                           (push (cons (+ num-var-cols col-offset) -1) cells)
                           (set-basis main row num-cols)
                           (incf col-offset))
                          (= (push row art-rows)
                           (set-basis main row num-cols))
                          (t (error 'parsing-error
                                    :description (format nil "~S is not a valid constraint equation"
                                                         constraint))))
                        (push (cons (1- num-cols) rhs) cells)
                        ;; term order, so that a later term overwrites an earlier one as (setf aref) does
                        (setf cells (nreverse cells)
                              (aref row-cells row) cells)
                        (loop for (c . v) in cells do (f64-set main row c v))))
             ;; objective row  :267-279
             (let ((obj-rhs 0))
               (loop for (var . coef) in (problem-objective-func problem)
                     for mp = (gethash var mappings)
                     do (ecase (first mp)
                          (:positive (f64-set main m (second mp) (- coef))
                           (incf obj-rhs (* coef (third mp))))
                          (:negative (f64-set main m (second mp) coef)
                           (incf obj-rhs (* coef (third mp))))
                          (:signed (f64-set main m (second mp) (- coef))
                           (f64-set main m (1+ (second mp)) coef))))
               (f64-set main m (1- num-cols) obj-rhs))
             ;; artificial tableau  :288-325
             (when art-rows
               (let* ((num-art (length art-rows))
                      (art-cols (+ num-cols num-art))
                      (sums (make-hash-table)))
                 (setf art (alloc-f64 (1+ m) art-cols))
                 (flet ((art-col (c) (if (= c (1- num-cols)) (1- art-cols) c)))
                   (dotimes (r m)
                     (set-basis art r (cffi:mem-aref (f64-tableau-basis main) :int32 r))
                     (loop for (c . v) in (aref row-cells r) do (f64-set art r (art-col c) v)))
                   (loop for row in art-rows
                         for i from 0
                         do (set-basis art row (+ num-cols -1 i))
                            (f64-set art row (+ num-cols -1 i) 1))
                   ;; phase-1 objective row: exact column sums over the artificial rows, ascending
                   (loop for row in (sort (copy-list art-rows) #'<)
                         do (loop for (c . v) in (aref row-cells row)
                                  do (incf (gethash c sums 0) v)))
                   (maphash (lambda (c v) (f64-set art m (art-col c) v)) sums))))
             (setf ok t)
             (values main art mappings))
        (unless ok (free-f64 main) (free-f64 art))))))

(defstruct (b200-solution (:constructor %make-b200-solution))
  "What the accessors of src/simplex.lisp:74-120 read, kept as flat double-float vectors."
  problem instance-problem mappings
  (var-count 0 :type fixnum) (constraint-count 0 :type fixnum)
  (rhs nil :type (simple-array double-float (*)))       ; column var-count, all rows
  (objective-row nil :type (simple-array double-float (*)))
  (basis nil :type (simple-array fixnum (*))))

(defun solution-basic-value (s column)
  (let ((idx (position column (b200-solution-basis s))))
    (if idx (aref (b200-solution-rhs s) idx) 0)))

(defmethod solution-problem ((s b200-solution)) (b200-solution-problem s))
(defmethod solution-objective-value ((s b200-solution))
  (aref (b200-solution-rhs s) (b200-solution-constraint-count s)))
(defmethod solution-variable ((s b200-solution) var)
  "tableau-variable, src/simplex.lisp:81-107"
  (if (eq var (problem-objective-var (b200-solution-instance-problem s)))
      (solution-objective-value s)
      (let ((mp (gethash var (b200-solution-mappings s))))
        (unless mp (error "~S is not a variable in the tableau" var))
        (ecase (first mp)
          (:positive (+ (third mp) (solution-basic-value s (second mp))))
          (:negative (+ (third mp) (- (solution-basic-value s (second mp)))))
          (:signed (- (solution-basic-value s (second mp))
                      (solution-basic-value s (1+ (second mp)))))))))
(defmethod solution-reduced-cost ((s b200-solution) var)
  "tableau-reduced-cost, src/simplex.lisp:111-120"
  (let ((mp (gethash var (b200-solution-mappings s))))
    (unless mp (error "~S is not a variable in the tableau" var))
    (unless (eq (first mp) :positive) (error "~S has no lower bound" var))
    (aref (b200-solution-objective-row s) (second mp))))

(defun gpu-solve-f64 (problem instance-problem &key (tolerance 1024) devices pivot-rule
                                                 max-iterations feas-mode)
  "build-tableau-f64 + b200lp_solve[_two_phase]; a B200-SOLUTION, or NIL when the problem has no
constraints (boxed path)."
  (multiple-value-bind (main art mappings) (build-tableau-f64 problem instance-problem)
    (unless main (return-from gpu-solve-f64 nil))
    (unwind-protect
         (let* ((rows (f64-tableau-rows main)) (cols (f64-tableau-cols main))
                (is-max (if (eq 'max (problem-type instance-problem)) 1 0)))
           (cffi:with-foreign-objects ((o '(:struct opts)) (res '(:struct result)))
             (fill-opts o tolerance devices pivot-rule max-iterations feas-mode)
             (signal-status
              (if art
                  (%solve-two-phase o (f64-tableau-tab art) (f64-tableau-cols art)
                                    (f64-tableau-cols art) (f64-tableau-basis art)
                                    (f64-tableau-tab main) rows cols cols (f64-tableau-basis main)
                                    is-max res)
                  (%solve o (f64-tableau-tab main) rows cols cols (f64-tableau-basis main)
                          is-max res (cffi:null-pointer) (cffi:null-pointer)))))
           (let ((rhs (make-array rows :element-type 'double-float))
                 (obj (make-array cols :element-type 'double-float))
                 (basis (make-array (1- rows) :element-type 'fixnum)))
             (dotimes (i rows)
               (setf (aref rhs i) (cffi:mem-aref (f64-tableau-tab main) :double
                                                 (+ (* i cols) (1- cols)))))
             (dotimes (j cols)
               (setf (aref obj j) (cffi:mem-aref (f64-tableau-tab main) :double
                                                 (+ (* (1- rows) cols) j))))
             (dotimes (i (1- rows))
               (setf (aref basis i) (cffi:mem-aref (f64-tableau-basis main) :int32 i)))
             (%make-b200-solution :problem problem :instance-problem instance-problem
                                  :mappings mappings :var-count (1- cols)
                                  :constraint-count (1- rows)
                                  :rhs rhs :objective-row obj :basis basis)))
      (free-f64 main)
      (free-f64 art))))

;;; branch and bound: control flow of simplex-solver (src/simplex.lisp:506-542) -------------------
(defun integral-p (value tolerance)
  "(integerp value) for exact rationals (src/simplex.lisp:479).  A double-float vertex with an
integer coordinate carries 1e-12..1e-10 of noise, so the test has a tolerance of its own, relative
to the value (never tighter than the fp= tolerance): branching on noise would add a row, a phase 1
and a GPU solve per spurious node."
  (<= (abs (- value (fround value)))
      (max (* 1d-9 (max 1d0 (abs value))) (* tolerance double-float-epsilon))))

(defun violated-integer-constraint (solution tolerance)
  "SOLUTION is a B200-SOLUTION or (boxed path) the reference's tableau: both answer the generics."
  (dolist (var (problem-integer-vars (solution-problem solution)))
    (unless (integral-p (solution-variable solution var) tolerance)
      (return var))))

(defun build-and-solve (problem extra tolerance boxed backend-args)
  (let ((instance (if (null extra)
                      problem
                      (linear-programming/problem::make-problem
                       :type (problem-type problem)
                       :vars (problem-vars problem)
                       :objective-var (problem-objective-var problem)
                       :objective-func (problem-objective-func problem)
                       :integer-vars (problem-integer-vars problem)
                       :var-bounds (problem-var-bounds problem)
                       :constraints (append extra (problem-constraints problem))))))
    (handler-case
        (or (and (not boxed) (apply #'gpu-solve-f64 problem instance backend-args))
            (apply #'gpu-solve-tableau
                   (build-tableau problem instance :fp-tolerance-factor tolerance)
                   backend-args))
      (infeasible-problem-error () :infeasible))))

(defun b200-solver (problem &key (fp-tolerance 1024) devices pivot-rule max-iterations feas-mode
                              boxed
                    &allow-other-keys)
  "The *solver* backend function.  Keywords: :fp-tolerance (as simplex-solver, src/simplex.lisp:511),
:devices (list of CUDA ordinals; more than one row-block shards the tableau), :pivot-rule
(0 reference rule, 1 Bland), :max-iterations, :feas-mode (0 scaled two-phase zero tests, 1 the
reference's literal ones; include/b200lp.h), :boxed (t: go through the reference's build-tableau
and return its tableau object instead of a B200-SOLUTION)."
  (let ((backend-args (list :tolerance fp-tolerance :devices devices :pivot-rule pivot-rule
                            :max-iterations max-iterations :feas-mode feas-mode))
        (better (if (eq (problem-type problem) 'max) #'< #'>))
        (best nil) (solution nil) (stack (list '())))
    (loop while stack
          do (let* ((entry (pop stack))
                    (tab (build-and-solve problem entry fp-tolerance boxed backend-args)))
               (unless (eq tab :infeasible)
                 (let ((violated (violated-integer-constraint tab fp-tolerance))
                       (value (solution-objective-value tab)))
                   (cond
                     ((and violated best (not (funcall better best value))))
                     (violated
                      (let ((val (solution-variable tab violated)))
                        (setf stack (list* (list* `(<= ((,violated . 1)) ,(floor val)) entry)
                                           (list* `(>= ((,violated . 1)) ,(ceiling val)) entry)
                                           stack))))
                     ((or (null best) (funcall better best value))
                      (setf best value solution tab)))))))
    (or solution (error 'infeasible-problem-error))))

(defun install ()
  "Make the B200 backend the default for solve-problem / with-solved-problem."
  (setf linear-programming/solver:*solver* 'b200-solver))
