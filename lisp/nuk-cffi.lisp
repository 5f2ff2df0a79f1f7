;;;; lisp/nuk-cffi.lisp -- CFFI stub for libnuk.so (NOT executed in this repository's CI: the build
;;;; image has no Common Lisp implementation).  It shows the exact reference-side binding a
;;;; maintainer would add; the same call sequence is exercised through Python ctypes in tests/.
;;;;
;;;; Level 0 (INTEGRATION.md): nothing but the library name changes -- cl-bmas keeps defining
;;;; bmas:sadd & co. with its own DEFCFUNs, which now resolve against libnuk.so.

(defpackage :nuk
  (:use :cl)
  (:export #:check #:malloc #:free #:h2d #:d2h #:map2 #:map1 #:cmp2 #:cast #:reduce
           #:+f32+ #:+f64+ #:+add+ #:+sub+ #:+mul+ #:+div+ #:+sum+ #:+hmax+ #:+hmin+))
(in-package :nuk)

(cffi:define-foreign-library libnuk (t (:or "libnuk.so")))
(cffi:use-foreign-library libnuk)

;;; dtype codes = column of the translation table (src/utils/translations.lisp:66-82)
(defconstant +f32+ 0) (defconstant +f64+ 1)
(defconstant +i64+ 3) (defconstant +i32+ 4) (defconstant +i16+ 5) (defconstant +i8+ 6)
(defconstant +u64+ 7) (defconstant +u32+ 8) (defconstant +u16+ 9) (defconstant +u8+ 10)
;;; op codes (include/nuk.h)
(defconstant +add+ 0) (defconstant +sub+ 1) (defconstant +mul+ 2) (defconstant +div+ 3)
(defconstant +max+ 4) (defconstant +min+ 5) (defconstant +pow+ 6) (defconstant +atan2+ 7)
(defconstant +logb+ 13) ; nu:log x y in one pass (replaces the log, log, divide of two-arg-fn-float.lisp:531-547)
(defconstant +sum+ 0) (defconstant +hmax+ 1) (defconstant +hmin+ 2)

(cffi:defcfun ("nuk_last_error" %last-error) :string)
(cffi:defcfun ("nuk_malloc" %malloc) :int (dptr :pointer) (bytes :size))
(cffi:defcfun ("nuk_free" %free) :int (dptr :pointer))
(cffi:defcfun ("nuk_memcpy_h2d" %h2d) :int (dst :pointer) (src :pointer) (bytes :size) (stream :pointer))
(cffi:defcfun ("nuk_memcpy_d2h" %d2h) :int (dst :pointer) (src :pointer) (bytes :size) (stream :pointer))
(cffi:defcfun ("nuk_stream_sync" %stream-sync) :int (stream :pointer))
(cffi:defcfun ("nuk_map2" %map2) :int
  (op :int) (dtype :int) (rank :int) (dims :pointer) (sx :pointer) (sy :pointer) (so :pointer)
  (x :pointer) (y :pointer) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_map1" %map1) :int
  (op :int) (dtype :int) (rank :int) (dims :pointer) (sx :pointer) (so :pointer)
  (x :pointer) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_cmp2" %cmp2) :int
  (cmp :int) (dtype :int) (rank :int) (dims :pointer) (sx :pointer) (sy :pointer) (so :pointer)
  (x :pointer) (y :pointer) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_cast" %cast) :int
  (src-dtype :int) (dst-dtype :int) (rank :int) (dims :pointer) (sx :pointer) (so :pointer)
  (x :pointer) (out :pointer) (stream :pointer))
(cffi:defcfun ("nuk_reduce" %reduce) :int
  (red :int) (dtype :int) (rank :int) (dims :pointer) (sx :pointer) (axis :int) (so :pointer)
  (x :pointer) (out :pointer) (flags :int) (stream :pointer))

(define-condition nuk-error (error)
  ((status :initarg :status) (message :initarg :message))
  (:report (lambda (c s) (format s "libnuk status ~D: ~A"
                                 (slot-value c 'status) (slot-value c 'message)))))

(declaim (inline check))
(defun check (status)
  (unless (zerop status)
    (error 'nuk-error :status status :message (%last-error)))
  status)

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

;;; One launch instead of the loop nest of src/utils/ptr-iterate-but-inner.lisp:35-64.
;;; X Y OUT are device pointers to element (0 ... 0); strides are ELEMENT strides of each
;;; operand over DIMS (0 on broadcast axes -- strides-for-broadcast, src/utils/broadcast.lisp:3-19
;;; -- or the view's own strides for dense-arrays).
(defun map2 (op dtype dims x sx y sy out so &optional (stream (cffi:null-pointer)))
  (with-int64-vectors ((d dims) (fx sx) (fy sy) (fo so))
    (check (%map2 op dtype (length dims) d fx fy fo x y out stream))))
