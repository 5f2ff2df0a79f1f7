;;;; lisp/nuk-cffi.lisp -- CFFI binding of libnuk.so, the host side the reference's maintainer would add.
;;;;
;;;; NOT executed in this repository (the build image has no Common Lisp implementation); every call
;;;; sequence below is exercised through Python ctypes in tests/ with the same arguments CFFI passes.
;;;; Three levels, as in INTEGRATION.md:
;;;;   0  nothing here is needed: cl-bmas loads libnuk.so instead of libbmas.so and its own DEFCFUNs of
;;;;      bmas:sadd & co. resolve against it (host vectors are staged through HBM by the library);
;;;;   1  device-resident storage for dense-arrays (DEVICE-STORAGE below) -- the polymorphs of
;;;;      src/basic-math, src/transcendental, src/statistics then run unmodified on device pointers;
;;;;   2  one launch per call: MAP1 MAP2 CMP2 CAST REDUCE FOLD MOMENTS ... replace the loop nests of
;;;;      src/utils/ptr-iterate-but-inner.lisp:35-64 and src/utils/simple-array-broadcast.lisp:7-84;
;;;;   +  multi-GPU: COMM-INIT-ALL / COMM-ALLREDUCE / COMM-REDUCE drive all GPUs of the box from one image.

(defpackage :nuk
  (:use :cl)
  (:shadow #:reduce #:fill)
  (:export #:nuk-error #:check
           ;; storage / streams
           #:init #:device-count #:malloc #:free #:h2d #:d2h #:d2d #:host-alloc #:host-free #:host-register
           #:host-unregister #:stream-create #:stream-destroy #:stream-sync #:set-stream #:device-sync
           #:set-pointer-mode #:+ptr-auto+ #:+ptr-device+
           #:device-storage #:make-device-storage #:device-storage-pointer #:free-device-storage
           #:upload-vector #:download-vector
           ;; graphs
           #:graph-begin #:graph-end #:graph-launch #:graph-destroy
           ;; one-launch nd entry points
           #:map1 #:map2 #:cmp2 #:cast #:map2-scalar #:cmp2-scalar #:fill #:fold #:reduce #:argreduce
           #:moments #:div-scalar-inplace #:variance-finish
           ;; multi-GPU
           #:shard-bounds #:shard-axis #:comm-init-all #:comm-destroy #:comm-status #:comm-allreduce
           #:comm-reduce #:comm-moments
           ;; codes
           #:+f32+ #:+f64+ #:+i64+ #:+i32+ #:+i16+ #:+i8+ #:+u64+ #:+u32+ #:+u16+ #:+u8+ #:dtype-code
           #:+add+ #:+sub+ #:+mul+ #:+div+ #:+max+ #:+min+ #:+pow+ #:+atan2+ #:+and+ #:+or+ #:+xor+ #:+andnot+
           #:+sra+ #:+logb+ #:+lt+ #:+le+ #:+eq+ #:+neq+ #:+ge+ #:+gt+
           #:+copy+ #:+abs+ #:+round+ #:+trunc+ #:+floor+ #:+ceil+ #:+sqrt+ #:+sin+ #:+cos+ #:+tan+ #:+asin+
           #:+acos+ #:+atan+ #:+sinh+ #:+cosh+ #:+tanh+ #:+asinh+ #:+acosh+ #:+atanh+ #:+exp+ #:+log+ #:+not+
           #:+sum+ #:+hmax+ #:+hmin+ #:+sumsq+ #:+reduce-fast+ #:+reduce-ref-order+))
(in-package :nuk)

(cffi:define-foreign-library libnuk (t (:or "libnuk.so")))
(cffi:use-foreign-library libnuk)

;;; ------------------------------------------------------------------ codes (include/nuk.h)
;;; dtype = column of the translation table (src/utils/translations.lisp:66-82)
(defconstant +f32+ 0) (defconstant +f64+ 1)
(defconstant +i64+ 3) (defconstant +i32+ 4) (defconstant +i16+ 5) (defconstant +i8+ 6)
(defconstant +u64+ 7) (defconstant +u32+ 8) (defconstant +u16+ 9) (defconstant +u8+ 10)
(defun dtype-code (element-type)
  "Lisp element type -> libnuk dtype code; NIL when there is no C type (translations.lisp:50-64)."
  (cond ((eq element-type 'single-float) +f32+)
        ((eq element-type 'double-float) +f64+)
        ((equal element-type '(signed-byte 64)) +i64+) ((equal element-type '(signed-byte 32)) +i32+)
        ((equal element-type '(signed-byte 16)) +i16+) ((equal element-type '(signed-byte 8)) +i8+)
        ((equal element-type '(unsigned-byte 64)) +u64+) ((equal element-type '(unsigned-byte 32)) +u32+)
        ((equal element-type '(unsigned-byte 16)) +u16+) ((equal element-type '(unsigned-byte 8)) +u8+)))
;; two-arg ops
(defconstant +add+ 0) (defconstant +sub+ 1) (defconstant +mul+ 2) (defconstant +div+ 3)
(defconstant +max+ 4) (defconstant +min+ 5) (defconstant +pow+ 6) (defconstant +atan2+ 7)
(defconstant +and+ 8) (defconstant +or+ 9) (defconstant +xor+ 10) (defconstant +andnot+ 11) (defconstant +sra+ 12)
(defconstant +logb+ 13) ; nu:log x y in one pass (replaces the log, log, divide of two-arg-fn-float.lisp:531-547)
;; comparisons
(defconstant +lt+ 0) (defconstant +le+ 1) (defconstant +eq+ 2) (defconstant +neq+ 3) (defconstant +ge+ 4) (defconstant +gt+ 5)
;; one-arg ops
(defconstant +copy+ 0) (defconstant +abs+ 1) (defconstant +round+ 2) (defconstant +trunc+ 3) (defconstant +floor+ 4)
(defconstant +ceil+ 5) (defconstant +sqrt+ 6) (defconstant +sin+ 7) (defconstant +cos+ 8) (defconstant +tan+ 9)
(defconstant +asin+ 10) (defconstant +acos+ 11) (defconstant +atan+ 12) (defconstant +sinh+ 13) (defconstant +cosh+ 14)
(defconstant +tanh+ 15) (defconstant +asinh+ 16) (defconstant +acosh+ 17) (defconstant +atanh+ 18) (defconstant +exp+ 19)
(defconstant +log+ 20) (defconstant +not+ 21)
;; reductions and their flags
(defconstant +sum+ 0) (defconstant +hmax+ 1) (defconstant +hmin+ 2) (defconstant +sumsq+ 3)
(defconstant +reduce-fast+ 0) (defconstant +reduce-ref-order+ 1)
(defconstant +ptr-auto+ 0) (defconstant +ptr-device+ 1)

;;; ------------------------------------------------------------------ raw entry points
(cffi:defcfun ("nuk_last_error" %last-error) :string)
(cffi:defcfun ("nuk_device_count" device-count) :int)
(cffi:defcfun ("nuk_init" %init) :int (device :int))
(cffi:defcfun ("nuk_malloc" %malloc) :int (dptr :pointer) (bytes :size))
(cffi:defcfun ("nuk_free" %free) :int (dptr :pointer))
(cffi:defcfun ("nuk_memcpy_h2d" %h2d) :int (dst :pointer) (src :pointer) (bytes :size) (stream :pointer))
(cffi:defcfun ("nuk_memcpy_d2h" %d2h) :int (dst :pointer) (src :pointer) (bytes :size) (stream :pointer))
(cffi:defcfun ("nuk_memcpy_d2d" %d2d) :int (dst :pointer) (src :pointer) (bytes :size) (stream :pointer))
(cffi:defcfun ("nuk_host_alloc" %host-alloc) :int (hptr :pointer) (bytes :size))
(cffi:defcfun ("nuk_host_free" %host-free) :int (hptr :pointer))
(cffi:defcfun ("nuk_host_register" %host-register) :int (hptr :pointer) (bytes :size))
(cffi:defcfun ("nuk_host_unregister" %host-unregister) :int (hptr :pointer))
(cffi:defcfun ("nuk_stream_create" %stream-create) :int (stream :pointer))
(cffi:defcfun ("nuk_stream_destroy" %stream-destroy) :int (stream :pointer))
(cffi:defcfun ("nuk_stream_sync" %stream-sync) :int (stream :pointer))
(cffi:defcfun ("nuk_set_stream" %set-stream) :int (stream :pointer))
(cffi:defcfun ("nuk_device_sync" %device-sync) :int)
(cffi:defcfun ("nuk_set_pointer_mode" %set-pointer-mode) :int (mode :int))
(cffi:defcfun ("nuk_graph_begin" %graph-begin) :int (stream :pointer))
(cffi:defcfun ("nuk_graph_end" %graph-end) :int (stream :pointer) (graph :pointer))
(cffi:defcfun ("nuk_graph_launch" %graph-launch) :int (graph :pointer) (stream :pointer))
(cffi:defcfun ("nuk_graph_destroy" %graph-destroy) :int (graph :pointer))
(cffi:defcfun ("nuk_map1" %map1) :int
  (op :int) (dtype :int) (rank :int) (dims :pointer) (sx :pointer) (so :pointer)
  (x :pointer) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_map2" %map2) :int
  (op :int) (dtype :int) (rank :int) (dims :pointer) (sx :pointer) (sy :pointer) (so :pointer)
  (x :pointer) (y :pointer) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_cmp2" %cmp2) :int
  (cmp :int) (dtype :int) (rank :int) (dims :pointer) (sx :pointer) (sy :pointer) (so :pointer)
  (x :pointer) (y :pointer) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_cast" %cast) :int
  (src-dtype :int) (dst-dtype :int) (rank :int) (dims :pointer) (sx :pointer) (so :pointer)
  (x :pointer) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_map2_scalar" %map2-scalar) :int
  (op :int) (dtype :int) (rank :int) (dims :pointer) (sa :pointer) (so :pointer)
  (a :pointer) (host-scalar :pointer) (scalar-is-x :int) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_cmp2_scalar" %cmp2-scalar) :int
  (cmp :int) (dtype :int) (rank :int) (dims :pointer) (sa :pointer) (so :pointer)
  (a :pointer) (host-scalar :pointer) (scalar-is-x :int) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_fill" %fill) :int
  (dtype :int) (rank :int) (dims :pointer) (so :pointer) (host-scalar :pointer) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_fold" %fold) :int
  (op :int) (dtype :int) (n :int64) (nin :int) (xs :pointer) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_reduce" %reduce) :int
  (red :int) (dtype :int) (rank :int) (dims :pointer) (sx :pointer) (axis :int) (so :pointer)
  (x :pointer) (out :pointer) (flags :int) (stream :pointer))
(cffi:defcfun ("nuk_argreduce" %argreduce) :int
  (is-max :int) (dtype :int) (rank :int) (dims :pointer) (sx :pointer) (axis :int) (so :pointer)
  (x :pointer) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_moments" %moments) :int
  (dtype :int) (rank :int) (dims :pointer) (sx :pointer) (axis :int) (so :pointer)
  (x :pointer) (sum-out :pointer) (sumsq-out :pointer) (flags :int) (stream :pointer))
(cffi:defcfun ("nuk_div_scalar_inplace" %div-scalar-inplace) :int
  (dtype :int) (n :int64) (out :pointer) (divisor :double) (stream :pointer))
(cffi:defcfun ("nuk_variance_finish" %variance-finish) :int
  (dtype :int) (n :int64) (s-sum :pointer) (q-sumsq :pointer) (count :double) (ddof :double) (stream :pointer))
(cffi:defcfun ("nuk_shard_bounds" %shard-bounds) :int (n :int64) (world :int) (rank :int) (lo :pointer) (hi :pointer))
(cffi:defcfun ("nuk_shard_axis" %shard-axis) :int (rank :int) (dims :pointer) (world :int))
(cffi:defcfun ("nuk_comm_init_all" %comm-init-all) :int (ndev :int) (devices :pointer) (comms :pointer))
(cffi:defcfun ("nuk_comm_destroy" %comm-destroy) :int (comm :pointer))
(cffi:defcfun ("nuk_comm_status" %comm-status) :int (comm :pointer))
(cffi:defcfun ("nuk_comm_allreduce" %comm-allreduce) :int
  (comm :pointer) (red :int) (dtype :int) (count :int64) (local :pointer) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_comm_reduce" %comm-reduce) :int
  (comm :pointer) (red :int) (dtype :int) (rank :int) (dims :pointer) (sx :pointer)
  (x :pointer) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_comm_moments" %comm-moments) :int
  (comm :pointer) (dtype :int) (rank :int) (dims :pointer) (sx :pointer)
  (x :pointer) (sum-out :pointer) (sumsq-out :pointer) (stream :pointer))

;;; ------------------------------------------------------------------ errors
(define-condition nuk-error (error)
  ((status :initarg :status :reader nuk-error-status) (message :initarg :message :reader nuk-error-message))
  (:report (lambda (c s) (format s "libnuk status ~D: ~A" (nuk-error-status c) (nuk-error-message c)))))

(declaim (inline check))
(defun check (status)
  "0 -> 0; anything else signals NUK-ERROR with nuk_last_error().  BMAS_* symbols return void / a value,
so after a bmas: call the Lisp side may call (check (nuk-last-status)) when it wants to know."
  (unless (zerop status)
    (error 'nuk-error :status status :message (%last-error)))
  status)

(defvar *stream* (cffi:null-pointer)
  "Stream argument of the nuk: entry points, in the style of nu:*multithreaded-threshold* & co.
(src/basic-utils/basic-utils.lisp:29-79).  The null pointer is the legacy default stream.")

(defmacro with-int64-vectors ((&rest bindings) &body body)
  "Each binding is (VAR LIST-OF-INTEGERS): a stack-allocated int64 array for a descriptor."
  (if (null bindings)
      `(progn ,@body)
      (destructuring-bind ((var list) &rest more) bindings
        (let ((l (gensym "LIST")))
          `(let ((,l ,list))
             (cffi:with-foreign-object (,var :int64 (max 1 (length ,l)))
               (loop :for i :from 0 :for d :in ,l
                     :do (setf (cffi:mem-aref ,var :int64 i) d))
               (with-int64-vectors ,more ,@body)))))))

(defmacro with-host-scalar ((var value dtype) &body body)
  "VALUE written to a one-element foreign buffer of DTYPE, the way the reference passes scalar operands
(src/basic-math/two-arg-fn-non-comparison.lisp:226-232)."
  (let ((v (gensym "V")) (dt (gensym "DT")))
    `(let ((,v ,value) (,dt ,dtype))
       (cffi:with-foreign-object (,var :uint64)
         (setf (cffi:mem-ref ,var :uint64) 0)
         (ecase ,dt
           (0 (setf (cffi:mem-ref ,var :float) (coerce ,v 'single-float)))
           (1 (setf (cffi:mem-ref ,var :double) (coerce ,v 'double-float)))
           (3 (setf (cffi:mem-ref ,var :int64) ,v)) (4 (setf (cffi:mem-ref ,var :int32) ,v))
           (5 (setf (cffi:mem-ref ,var :int16) ,v)) (6 (setf (cffi:mem-ref ,var :int8) ,v))
           (7 (setf (cffi:mem-ref ,var :uint64) ,v)) (8 (setf (cffi:mem-ref ,var :uint32) ,v))
           (9 (setf (cffi:mem-ref ,var :uint16) ,v)) (10 (setf (cffi:mem-ref ,var :uint8) ,v)))
         ,@body))))

;;; ------------------------------------------------------------------ runtime, storage (Level 1)
(defun init (&optional (device 0)) (check (%init device)))
(defun malloc (bytes)
  (cffi:with-foreign-object (p :pointer)
    (check (%malloc p bytes))
    (cffi:mem-ref p :pointer)))
(defun free (dptr) (check (%free dptr)))
(defun h2d (dst src bytes &optional (stream *stream*)) (check (%h2d dst src bytes stream)))
(defun d2h (dst src bytes &optional (stream *stream*)) (check (%d2h dst src bytes stream)))
(defun d2d (dst src bytes &optional (stream *stream*)) (check (%d2d dst src bytes stream)))
(defun host-alloc (bytes)
  (cffi:with-foreign-object (p :pointer)
    (check (%host-alloc p bytes))
    (cffi:mem-ref p :pointer)))
(defun host-free (hptr) (check (%host-free hptr)))
(defun host-register (hptr bytes) (check (%host-register hptr bytes)))
(defun host-unregister (hptr) (check (%host-unregister hptr)))
(defun stream-create ()
  (cffi:with-foreign-object (p :pointer)
    (check (%stream-create p))
    (cffi:mem-ref p :pointer)))
(defun stream-destroy (stream) (check (%stream-destroy stream)))
(defun stream-sync (&optional (stream *stream*)) (check (%stream-sync stream)))
(defun set-stream (stream) (check (%set-stream stream)))   ; the stream bmas: calls of this thread run on
(defun device-sync () (check (%device-sync)))
(defun set-pointer-mode (mode)
  "+PTR-DEVICE+ once every array handed to bmas: lives in DEVICE-STORAGE: no driver query per operand."
  (check (%set-pointer-mode mode)))

(defstruct (device-storage (:constructor %make-device-storage))
  "Storage of a dense-arrays array in HBM: what array-storage returns for the device backend
(replaces the Lisp heap vector of dense-numericals-src/utils/primitives.lisp:3-57)."
  (pointer (cffi:null-pointer) :type cffi:foreign-pointer)
  (element-type 'single-float)
  (size 0 :type (unsigned-byte 62)))

(defun make-device-storage (size element-type)
  (%make-device-storage :pointer (malloc (* size (cffi:foreign-type-size (c-type element-type))))
                        :element-type element-type :size size))
(defun free-device-storage (storage)
  (free (device-storage-pointer storage))
  (setf (device-storage-pointer storage) (cffi:null-pointer)))
(defun c-type (element-type)
  (ecase (dtype-code element-type)
    (0 :float) (1 :double) (3 :int64) (4 :int32) (5 :int16) (6 :int8) (7 :uint64) (8 :uint32) (9 :uint16) (10 :uint8)))

(defun upload-vector (vector storage)
  "Lisp (simple-array ELEMENT-TYPE (*)) -> HBM; the vector is pinned against the GC for the copy, the way
with-pointers-to-vectors-data does it (src/basic-utils/common.lisp:207-212)."
  (cffi:with-pointer-to-vector-data (p vector)
    (h2d (device-storage-pointer storage) p
         (* (length vector) (cffi:foreign-type-size (c-type (device-storage-element-type storage)))))
    (stream-sync))
  storage)
(defun download-vector (storage vector)
  (cffi:with-pointer-to-vector-data (p vector)
    (d2h p (device-storage-pointer storage)
         (* (length vector) (cffi:foreign-type-size (c-type (device-storage-element-type storage)))))
    (stream-sync))
  vector)

;;; ------------------------------------------------------------------ CUDA graphs (launch-bound loops, C1)
(defun graph-begin (stream) (check (%graph-begin stream)))
(defun graph-end (stream)
  (cffi:with-foreign-object (g :pointer)
    (check (%graph-end stream g))
    (cffi:mem-ref g :pointer)))
(defun graph-launch (graph stream) (check (%graph-launch graph stream)))
(defun graph-destroy (graph) (check (%graph-destroy graph)))

;;; ------------------------------------------------------------------ one launch per call (Level 2)
;;; X Y OUT are device pointers to element (0 ... 0); strides are ELEMENT strides of each operand over DIMS
;;; (0 on broadcast axes -- strides-for-broadcast, src/utils/broadcast.lisp:3-19 -- or the view's own
;;; strides for dense-arrays).  Each function replaces one expansion of ptr-iterate-but-inner.
(defun map1 (op dtype dims x sx out so &optional (stream *stream*))
  "one-arg-fn/float, one-arg-fn/all and the transcendental unaries (src/basic-math/one-arg-fn-float.lisp:97-144)"
  (with-int64-vectors ((d dims) (fx sx) (fo so))
    (check (%map1 op dtype (length dims) d fx fo x out stream))))
(defun map2 (op dtype dims x sx y sy out so &optional (stream *stream*))
  "two-arg-fn/non-comparison, two-arg-fn/float (src/basic-math/two-arg-fn-non-comparison.lisp:151-186)"
  (with-int64-vectors ((d dims) (fx sx) (fy sy) (fo so))
    (check (%map2 op dtype (length dims) d fx fy fo x y out stream))))
(defun cmp2 (cmp dtype dims x sx y sy out so &optional (stream *stream*))
  "two-arg-fn/comparison: OUT is (unsigned-byte 8) 0/1 (src/basic-math/two-arg-fn-comparison.lisp:142-177)"
  (with-int64-vectors ((d dims) (fx sx) (fy sy) (fo so))
    (check (%cmp2 cmp dtype (length dims) d fx fy fo x y out stream))))
(defun cast (src-dtype dst-dtype dims x sx out so &optional (stream *stream*))
  "nu:copy / nu:astype (src/basic-math/copy-coerce.lisp:8-201)"
  (with-int64-vectors ((d dims) (fx sx) (fo so))
    (check (%cast src-dtype dst-dtype (length dims) d fx fo x out stream))))
(defun map2-scalar (op dtype dims a sa scalar scalar-is-x out so &optional (stream *stream*))
  "array-scalar polymorphs (src/basic-math/two-arg-fn-non-comparison.lisp:214-281): the scalar travels by value"
  (with-int64-vectors ((d dims) (fa sa) (fo so))
    (with-host-scalar (s scalar dtype)
      (check (%map2-scalar op dtype (length dims) d fa fo a s (if scalar-is-x 1 0) out stream)))))
(defun cmp2-scalar (cmp dtype dims a sa scalar scalar-is-x out so &optional (stream *stream*))
  (with-int64-vectors ((d dims) (fa sa) (fo so))
    (with-host-scalar (s scalar dtype)
      (check (%cmp2-scalar cmp dtype (length dims) d fa fo a s (if scalar-is-x 1 0) out stream)))))
(defun fill (dtype dims value out so &optional (stream *stream*))
  (with-int64-vectors ((d dims) (fo so))
    (with-host-scalar (s value dtype)
      (check (%fill dtype (length dims) d fo s out stream)))))
(defun fold (op dtype n operands out &optional (stream *stream*))
  "n-ary nu:+ nu:- nu:* nu:/ nu:max nu:min over 2..8 contiguous arrays in ONE pass
(src/basic-math/n-arg-fn.lisp:59-88 makes k-1 passes over OUT); same left-to-right association."
  (cffi:with-foreign-object (xs :pointer (length operands))
    (loop :for i :from 0 :for p :in operands :do (setf (cffi:mem-aref xs :pointer i) p))
    (check (%fold op dtype n (length operands) xs out stream))))
(defun reduce (red dtype dims x sx axis out so &key (flags +reduce-fast+) (stream *stream*))
  "nu:sum / nu:maximum / nu:minimum (src/basic-math/sum.lisp:52-127): AXIS -1 = everything; SO are OUT's
strides over DIMS without AXIS.  +REDUCE-REF-ORDER+ = the reference's row-accumulate order bit for bit."
  (with-int64-vectors ((d dims) (fx sx) (fo so))
    (check (%reduce red dtype (length dims) d fx axis fo x out flags stream))))
(defun argreduce (is-max dtype dims x sx axis out so &optional (stream *stream*))
  "nu:arg-maximum / nu:arg-minimum along AXIS (src/basic-math/arg-maximum.lisp:31-73); OUT is (signed-byte 64)"
  (with-int64-vectors ((d dims) (fx sx) (fo so))
    (check (%argreduce (if is-max 1 0) dtype (length dims) d fx axis fo x out stream))))
(defun moments (dtype dims x sx axis sum-out sumsq-out so &key (flags +reduce-fast+) (stream *stream*))
  "sum x and sum x^2 in one pass: the two sums of nu:variance (src/statistics/variance.lisp:55-68)"
  (with-int64-vectors ((d dims) (fx sx) (fo so))
    (check (%moments dtype (length dims) d fx axis fo x sum-out sumsq-out flags stream))))
(defun div-scalar-inplace (dtype n out divisor &optional (stream *stream*))
  "the tail of nu:mean (src/statistics/mean.lisp:90-97)"
  (check (%div-scalar-inplace dtype n out (coerce divisor 'double-float) stream)))
(defun variance-finish (dtype n s-sum q-sumsq count ddof &optional (stream *stream*))
  "the rounded tail of nu:variance (src/statistics/variance.lisp:58-68); the result is left in Q-SUMSQ"
  (check (%variance-finish dtype n s-sum q-sumsq (coerce count 'double-float) (coerce ddof 'double-float) stream)))

;;; ------------------------------------------------------------------ multi-GPU, one image driving all devices
(defun shard-bounds (n world rank)
  "slab [lo, hi) of RANK: the reference's max-work-size rule (src/transcendental/lparallel.lisp:66-81)"
  (cffi:with-foreign-objects ((lo :int64) (hi :int64))
    (check (%shard-bounds n world rank lo hi))
    (values (cffi:mem-ref lo :int64) (cffi:mem-ref hi :int64))))
(defun shard-axis (dims world)
  "first axis long enough for WORLD workers (dense-numericals-src/utils/lparallel.lisp:41-49), or NIL"
  (with-int64-vectors ((d dims))
    (let ((a (%shard-axis (length dims) d world)))
      (if (minusp a) nil a))))
(defun comm-init-all (ndev)
  "One communicator per device 0 .. NDEV-1 (peer access enabled both ways).  Returns a list."
  (cffi:with-foreign-object (cs :pointer ndev)
    (check (%comm-init-all ndev (cffi:null-pointer) cs))
    (loop :for i :below ndev :collect (cffi:mem-aref cs :pointer i))))
(defun comm-destroy (comm) (check (%comm-destroy comm)))
(defun comm-status (comm) (check (%comm-status comm)))
(defun comm-allreduce (comms red dtype count locals outs &optional (streams nil))
  "Rank-ordered all-reduce of COUNT elements: LOCALS / OUTS are per-device pointers, one per communicator.
All launches are issued before anything is waited for (each call is asynchronous and switches device itself)."
  (loop :for c :in comms :for l :in locals :for o :in outs :for i :from 0
        :do (check (%comm-allreduce c red dtype count l o (if streams (nth i streams) (cffi:null-pointer))))))
(defun comm-reduce (comms red dtype slab-dims slab-strides xs outs &optional (streams nil))
  "FULL reduction of an array sharded in slabs: SLAB-DIMS / SLAB-STRIDES / XS / OUTS are per-device lists.
Every OUT receives the same bits (rank-ordered fold inside the last kernel)."
  (loop :for c :in comms :for dims :in slab-dims :for sx :in slab-strides :for x :in xs :for o :in outs :for i :from 0
        :do (with-int64-vectors ((d dims) (fx sx))
              (check (%comm-reduce c red dtype (length dims) d fx x o
                                   (if streams (nth i streams) (cffi:null-pointer)))))))
(defun comm-moments (comms dtype slab-dims slab-strides xs sum-outs sumsq-outs &optional (streams nil))
  (loop :for c :in comms :for dims :in slab-dims :for sx :in slab-strides :for x :in xs
        :for s :in sum-outs :for q :in sumsq-outs :for i :from 0
        :do (with-int64-vectors ((d dims) (fx sx))
              (check (%comm-moments c dtype (length dims) d fx x s q
                                    (if streams (nth i streams) (cffi:null-pointer)))))))

;;; ------------------------------------------------------------------ example: (nu:add a b :out c) on device storage
;;; (let* ((n 1000000)
;;;        (a (make-device-storage n 'double-float)) (b (make-device-storage n 'double-float))
;;;        (c (make-device-storage n 'double-float)))
;;;   (upload-vector host-a a) (upload-vector host-b b)
;;;   ;; Level 1: the reference's own polymorph body, unchanged -- bmas:dadd now gets device pointers
;;;   (bmas:dadd n (device-storage-pointer a) 1 (device-storage-pointer b) 1 (device-storage-pointer c) 1)
;;;   ;; Level 2: any broadcast / strided shape in one launch
;;;   (map2 +add+ +f64+ '(1000 1000) (device-storage-pointer a) '(1000 1) (device-storage-pointer b) '(0 1)
;;;         (device-storage-pointer c) '(1000 1))
;;;   (download-vector c host-c))
