(* nx-cuda: Nx_backend over libnxcuda.so (include/nxcuda.h).

   The shape of this file follows the reference's C-backend veneer
   (packages/nx/lib/backend_c/nx_backend.ml): compute ops allocate a C-contiguous
   output and pass (out, inputs...) to one external; movement ops are View rewrites
   sharing the buffer; reduce drops the reduced axes; argmax/argmin honour keepdims.
   UNVERIFIED: never compiled (no OCaml toolchain in the build image). *)

open Nx_core

type device_buffer (* custom block: { void *ptr; size_t bytes; nxc_ctx *ctx } *)
type context (* custom block around nxc_ctx* *)

external create_context : unit -> context = "nx_cuda_ctx_create"
external dev_alloc : context -> int -> device_buffer = "nx_cuda_alloc"
external dev_of_host : context -> ('a, 'b) Nx_buffer.t -> device_buffer = "nx_cuda_of_host"
external dev_to_host : context -> device_buffer -> ('a, 'b) Nx_buffer.t -> unit = "nx_cuda_to_host"

(* FIELD ORDER IS ABI (slots 0-3 are read by the stubs; slot 4 carries Dtype.Packed.tag). *)
type ('a, 'b) t = {
  buffer : device_buffer;
  shape : int array;
  strides : int array;
  offset : int;
  tag : int;
  context : context;
  dtype : ('a, 'b) Dtype.t;
  elems : int; (* elements in the underlying buffer *)
}

let view t = View.create ~offset:t.offset ~strides:t.strides t.shape
let dtype t = t.dtype
let context t = t.context

let create_tensor ctx (dtype : ('a, 'b) Dtype.t) shape : ('a, 'b) t =
  let n = Array.fold_left ( * ) 1 shape in
  let bytes = max 16 (n * Dtype.itemsize dtype) in
  { buffer = dev_alloc ctx bytes; shape; strides = Shape.c_contiguous_strides shape; offset = 0;
    tag = Dtype.Packed.tag (Dtype.Packed.pack dtype); context = ctx; dtype; elems = n }

let buffer ctx dtype shape = create_tensor ctx dtype shape

external caml_fill : ('a, 'b) t -> ('a, 'b) Nx_buffer.t -> unit = "nx_cuda_fill"

let full ctx dtype shape value =
  let t = create_tensor ctx dtype shape in
  let one = Nx_buffer.create dtype 1 in
  Nx_buffer.set one 0 value;          (* typed store; the byte pattern travels as a kernel argument *)
  caml_fill t one;
  t

let from_host ctx buf =
  let dtype = Nx_buffer.kind buf and n = Nx_buffer.length buf in
  { buffer = dev_of_host ctx buf; shape = [| n |]; strides = [| 1 |]; offset = 0;
    tag = Dtype.Packed.tag (Dtype.Packed.pack dtype); context = ctx; dtype; elems = n }

(* Device backends copy the storage out (backend_intf.ml:108-116); the frontend indexes the
   result with the view's offset/strides (frontend.ml:1703-1708). *)
let to_host t =
  let host = Nx_buffer.create t.dtype t.elems in
  dev_to_host t.context t.buffer host;
  host

let of_view t v = { t with shape = View.shape v; strides = View.strides v; offset = View.offset v }
let expand t shape = of_view t (View.expand (view t) shape)
let reshape t shape = of_view t (View.reshape (view t) shape)
let permute t axes = of_view t (View.permute (view t) axes)
let shrink t bounds = of_view t (View.shrink (view t) bounds)
let flip t axes = of_view t (View.flip (view t) axes)
let is_c_contiguous t = View.is_c_contiguous (view t) && t.offset = 0

(* ---- map family: op codes are nxc_map1_op / nxc_map2_op / nxc_cmp_op ---- *)
external caml_map1 : int -> ('a, 'b) t -> ('a, 'b) t -> unit = "nx_cuda_map1"
external caml_map2 : int -> ('a, 'b) t -> ('a, 'b) t -> ('a, 'b) t -> unit = "nx_cuda_map2"
external caml_cmp : int -> (bool, Dtype.bool_elt) t -> ('a, 'b) t -> ('a, 'b) t -> unit = "nx_cuda_cmp"
external caml_where : ('a, 'b) t -> (bool, Dtype.bool_elt) t -> ('a, 'b) t -> ('a, 'b) t -> unit = "nx_cuda_where"
external caml_cast : ('c, 'd) t -> ('a, 'b) t -> unit = "nx_cuda_cast"
external caml_copy : ('a, 'b) t -> ('a, 'b) t -> unit = "nx_cuda_copy"

let unary op x = let out = create_tensor x.context x.dtype x.shape in caml_map1 op out x; out
let binary op x y = let out = create_tensor x.context x.dtype x.shape in caml_map2 op out x y; out
let comparison op x y = let out = create_tensor x.context Dtype.Bool x.shape in caml_cmp op out x y; out

let neg x = unary 0 x and recip x = unary 1 x and abs x = unary 2 x and sign x = unary 3 x
let sqrt x = unary 4 x and exp x = unary 5 x and log x = unary 6 x and sin x = unary 7 x
let cos x = unary 8 x and tan x = unary 9 x and asin x = unary 10 x and acos x = unary 11 x
let atan x = unary 12 x and sinh x = unary 13 x and cosh x = unary 14 x and tanh x = unary 15 x
let trunc x = unary 16 x and ceil x = unary 17 x and floor x = unary 18 x and round x = unary 19 x
let erf x = unary 20 x
let add x y = binary 0 x y and sub x y = binary 1 x y and mul x y = binary 2 x y
let idiv x y = binary 3 x y and fdiv x y = binary 4 x y and mod_ x y = binary 5 x y
let max x y = binary 6 x y and min x y = binary 7 x y and pow x y = binary 8 x y
let atan2 x y = binary 9 x y and xor x y = binary 10 x y and or_ x y = binary 11 x y
let and_ x y = binary 12 x y
let cmpeq x y = comparison 0 x y and cmpne x y = comparison 1 x y
let cmplt x y = comparison 2 x y and cmple x y = comparison 3 x y

let where cond a b = let out = create_tensor a.context a.dtype a.shape in caml_where out cond a b; out
let cast ~dtype x = let out = create_tensor x.context dtype x.shape in caml_cast out x; out
let copy x = let out = create_tensor x.context x.dtype x.shape in caml_copy out x; out
let contiguous x = if is_c_contiguous x then x else copy x
let assign dst src = caml_copy dst src

(* ---- fold family ---- *)
external caml_reduce : int -> ('a, 'b) t -> ('a, 'b) t -> int array -> unit = "nx_cuda_reduce"
external caml_argreduce : bool -> (int32, Dtype.int32_elt) t -> ('a, 'b) t -> int -> unit = "nx_cuda_argreduce"
external caml_scan : int -> ('a, 'b) t -> ('a, 'b) t -> int -> unit = "nx_cuda_scan"

let reduce ~op ~axes x =
  let code, extreme =
    match op with `Sum -> (0, None) | `Prod -> (1, None) | `Max -> (2, Some "reduce_max") | `Min -> (3, Some "reduce_min") in
  let axes = Array.copy axes in
  Array.sort Stdlib.compare axes;
  (match extreme with
   | Some name ->
       Array.iter (fun ax -> if x.shape.(ax) = 0 then
         invalid_arg (name ^ ": reduction over an empty axis has no identity")) axes
   | None -> ());
  let out = create_tensor x.context x.dtype (Shape.reduce_output_shape x.shape axes false) in
  caml_reduce code out x axes;
  out

let argreduce name is_max ~axis ~keepdims x =
  if x.shape.(axis) = 0 then invalid_arg (name ^ ": argument reduction over an empty axis");
  let out = create_tensor x.context Dtype.Int32 (Shape.reduce_output_shape x.shape [| axis |] keepdims) in
  caml_argreduce is_max out x axis;
  out

let argmax ~axis ~keepdims x = argreduce "argmax" true ~axis ~keepdims x
let argmin ~axis ~keepdims x = argreduce "argmin" false ~axis ~keepdims x

let associative_scan ~axis ~op x =
  let code = match op with `Sum -> 0 | `Prod -> 1 | `Max -> 2 | `Min -> 3 in
  let out = create_tensor x.context x.dtype x.shape in
  caml_scan code out x axis;
  out

(* ---- move / index / random ---- *)
external caml_pad : ('a, 'b) t -> ('a, 'b) t -> ('a, 'b) Nx_buffer.t -> int array -> unit = "nx_cuda_pad"
external caml_cat : ('a, 'b) t -> ('a, 'b) t array -> int -> unit = "nx_cuda_cat"
external caml_gather : ('a, 'b) t -> ('a, 'b) t -> (int32, Dtype.int32_elt) t -> int -> unit = "nx_cuda_gather"
external caml_scatter : ('a, 'b) t -> (int32, Dtype.int32_elt) t -> ('a, 'b) t -> int -> int -> unit = "nx_cuda_scatter"
external caml_threefry : (int32, Dtype.int32_elt) t -> (int32, Dtype.int32_elt) t -> (int32, Dtype.int32_elt) t -> unit = "nx_cuda_threefry"

let pad x padding fill_value =
  let out_shape = Array.mapi (fun i d -> let b, a = padding.(i) in d + b + a) x.shape in
  let out = create_tensor x.context x.dtype out_shape in
  let one = Nx_buffer.create x.dtype 1 in
  Nx_buffer.set one 0 fill_value;
  caml_pad out x one (Array.map fst padding);
  out

let cat tensors ~axis =
  match tensors with
  | [] -> invalid_arg "cat: empty tensor list"
  | first :: _ ->
      let ndim = Array.length first.shape in
      let axis = if axis < 0 then axis + ndim else axis in
      let total = List.fold_left (fun acc t -> acc + t.shape.(axis)) 0 tensors in
      let out_shape = Array.mapi (fun i d -> if i = axis then total else d) first.shape in
      let out = create_tensor first.context first.dtype out_shape in
      caml_cat out (Array.of_list tensors) axis;
      out

let gather data indices ~axis =
  let out = create_tensor data.context data.dtype indices.shape in
  caml_gather out data indices axis;
  out

let scatter ~mode ~unique_indices template ~indices ~updates ~axis =
  let out = copy template in
  let m = (match mode with `Set -> 0 | `Add -> 1) lor (if unique_indices then 2 else 0) in
  caml_scatter out indices updates axis m;
  out

let threefry key counter =
  let out = create_tensor counter.context Dtype.Int32 counter.shape in
  caml_threefry out key counter;
  out

(* ---- sort family ---- *)
external caml_sort : bool -> ('c, 'd) t -> ('a, 'b) t -> int -> bool -> unit = "nx_cuda_sort"

let sort ~axis ~descending x =
  let out = create_tensor x.context x.dtype x.shape in
  caml_sort false out x axis descending;
  out

let argsort ~axis ~descending x =
  let out = create_tensor x.context Dtype.Int32 x.shape in
  caml_sort true out x axis descending;
  out

(* ---- window ops (im2col / col2im) ---- *)
external caml_unfold : ('a, 'b) t -> ('a, 'b) t -> int array -> int array -> int array -> int array -> unit
  = "nx_cuda_unfold_bc" "nx_cuda_unfold"
external caml_fold : ('a, 'b) t -> ('a, 'b) t -> int array -> int array -> int array -> int array -> int array -> unit
  = "nx_cuda_fold_bc" "nx_cuda_fold"

let flat_pairs p = Array.concat (Array.to_list (Array.map (fun (b, a) -> [| b; a |]) p))

let unfold x ~kernel_size ~stride ~dilation ~padding =
  let k = Array.length kernel_size in
  let nlead = Array.length x.shape - k in
  let windows i =
    let before, after = padding.(i) in
    ((x.shape.(nlead + i) + before + after - (dilation.(i) * (kernel_size.(i) - 1) + 1)) / stride.(i)) + 1 in
  let l = Array.fold_left ( * ) 1 (Array.init k windows) in
  let kp = Array.fold_left ( * ) 1 kernel_size in
  let out = create_tensor x.context x.dtype (Array.append (Array.sub x.shape 0 nlead) [| kp; l |]) in
  caml_unfold out x kernel_size stride dilation (flat_pairs padding);
  out

let fold x ~output_size ~kernel_size ~stride ~dilation ~padding =
  let nlead = Array.length x.shape - 2 in
  let out = create_tensor x.context x.dtype (Array.append (Array.sub x.shape 0 nlead) output_size) in
  caml_fold out x output_size kernel_size stride dilation (flat_pairs padding);
  out

(* ---- matmul ---- *)
external caml_matmul : ('a, 'b) t -> ('a, 'b) t -> ('a, 'b) t -> unit = "nx_cuda_matmul"

let matmul x y =
  let xs = x.shape and ys = y.shape in
  let xnd = Array.length xs and ynd = Array.length ys in
  let m = xs.(xnd - 2) and n = ys.(ynd - 1) in
  let max_nd = Int.max xnd ynd in
  let batch = Array.init (max_nd - 2) (fun i ->
    let ai = i - (max_nd - xnd) and bi = i - (max_nd - ynd) in
    Int.max (if ai >= 0 then xs.(ai) else 1) (if bi >= 0 then ys.(bi) else 1)) in
  let out = create_tensor x.context x.dtype (Array.append batch [| m; n |]) in
  caml_matmul out x y;
  out

(* ---- fft family (backend_c/nx_backend.ml:502-549): unnormalised; the binding owns the
   output shape, the engine reads only the last entry of [s] ---- *)
external caml_fft : bool -> (Complex.t, 'b) t -> (Complex.t, 'b) t -> int array -> unit = "nx_cuda_fft"
external caml_rfft : (Complex.t, 'b) t -> (float, 'a) t -> int array -> unit = "nx_cuda_rfft"
external caml_irfft : (float, 'b) t -> (Complex.t, 'a) t -> int array -> int array -> unit = "nx_cuda_irfft"

let fft x ~axes =
  let out = create_tensor x.context x.dtype x.shape in
  caml_fft false out x axes;
  out

let ifft x ~axes =
  let out = create_tensor x.context x.dtype x.shape in
  caml_fft true out x axes;
  out

let rfft x ~dtype ~axes =
  let last = axes.(Array.length axes - 1) in
  let out_shape = Array.copy x.shape in
  out_shape.(last) <- (x.shape.(last) / 2) + 1;
  let out = create_tensor x.context dtype out_shape in
  caml_rfft out x axes;
  out

let irfft ?s x ~dtype ~axes =
  let last_idx = Array.length axes - 1 in
  let last = axes.(last_idx) in
  let size = match s with Some sizes -> sizes.(last_idx) | None -> (x.shape.(last) - 1) * 2 in
  let out_shape = Array.copy x.shape in
  out_shape.(last) <- size;
  let out = create_tensor x.context dtype out_shape in
  caml_irfft out x axes (match s with Some sizes -> sizes | None -> [||]);
  out

(* ---- linalg tier 1 (backend_c/nx_backend.ml:551-625). Numeric failures arrive as
   [Failure "<op>: <reason>"] and are lifted to [Linalg_error] by suffix, like the reference's
   [reraise_linalg]; the three booleans of the solve travel as one int (bit 0 upper, 1 transpose,
   2 unit diagonal). ---- *)
external caml_cholesky : ('a, 'b) t -> ('a, 'b) t -> bool -> unit = "nx_cuda_cholesky"
external caml_trsm : ('a, 'b) t -> ('a, 'b) t -> ('a, 'b) t -> int -> unit = "nx_cuda_triangular_solve"
external caml_qr : ('a, 'b) t -> ('a, 'b) t -> ('a, 'b) t -> bool -> unit = "nx_cuda_qr"

let lift_linalg ~op f =
  try f () with Failure msg as e ->
    let kind =
      if String.ends_with ~suffix:"matrix is not positive definite" msg then Some `Not_positive_definite
      else if String.ends_with ~suffix:"triangular matrix is singular" msg then Some `Singular
      else if String.ends_with ~suffix:"eigenvalue iteration did not converge" msg then Some `No_convergence
      else None in
    (match kind with Some kind -> raise (Backend_intf.Linalg_error { op; kind }) | None -> raise e)

let cholesky ~upper x =
  let out = create_tensor x.context x.dtype x.shape in
  lift_linalg ~op:"cholesky" (fun () -> caml_cholesky out x upper);
  out

let triangular_solve ~upper ~transpose ~unit_diag a b =
  let vec = Array.length b.shape = Array.length a.shape - 1 in
  let bm = if vec then reshape b (Array.append b.shape [| 1 |]) else b in
  let out = create_tensor b.context b.dtype bm.shape in
  let flags = Bool.to_int upper lor (Bool.to_int transpose lsl 1) lor (Bool.to_int unit_diag lsl 2) in
  lift_linalg ~op:"triangular_solve" (fun () -> caml_trsm out a bm flags);
  if vec then reshape out b.shape else out

let qr ~reduced x =
  let nd = Array.length x.shape in
  let m = x.shape.(nd - 2) and n = x.shape.(nd - 1) in
  let k = Int.min m n in
  let qs = Array.copy x.shape and rs = Array.copy x.shape in
  if reduced then (qs.(nd - 1) <- k; rs.(nd - 2) <- k) else qs.(nd - 1) <- m;
  let q = create_tensor x.context x.dtype qs and r = create_tensor x.context x.dtype rs in
  lift_linalg ~op:"qr" (fun () -> caml_qr q r x reduced);
  (q, r)

(* ---- linalg tier 2 (backend_c/nx_backend.ml:627-648): eigenvalues are always float64; the
   values-only call passes the input in the eigenvector slot, which the engine then ignores ---- *)
external caml_eigh : (float, Dtype.float64_elt) t -> ('a, 'b) t -> ('a, 'b) t -> bool -> unit = "nx_cuda_eigh"

let eigh_values x =
  let nd = Array.length x.shape in
  create_tensor x.context Dtype.Float64 (Array.append (Array.sub x.shape 0 (nd - 2)) [| x.shape.(nd - 1) |])

let eigvalsh x =
  let w = eigh_values x in
  lift_linalg ~op:"eigvalsh" (fun () -> caml_eigh w x x false);
  w

let eigh x =
  let w = eigh_values x and v = create_tensor x.context x.dtype x.shape in
  lift_linalg ~op:"eigh" (fun () -> caml_eigh w v x true);
  (w, v)

(* ---- linalg tier 3 (backend_c/nx_backend.ml:650-707): S is always float64 and the thin / full
   choice travels in the U / V^H shapes; eig's outputs are always complex128 and the values-only
   call passes [w] in the eigenvector slot, which the engine then ignores ---- *)
external caml_svd : ('a, 'b) t -> (float, Dtype.float64_elt) t -> ('a, 'b) t -> ('a, 'b) t -> unit = "nx_cuda_svd"

let svd ~full_matrices x =
  let nd = Array.length x.shape in
  let m = x.shape.(nd - 2) and n = x.shape.(nd - 1) in
  let k = Int.min m n in
  let batch = Array.sub x.shape 0 (nd - 2) in
  let u = create_tensor x.context x.dtype (Array.append batch (if full_matrices then [| m; m |] else [| m; k |])) in
  let s = create_tensor x.context Dtype.Float64 (Array.append batch [| k |]) in
  let vt = create_tensor x.context x.dtype (Array.append batch (if full_matrices then [| n; n |] else [| k; n |])) in
  lift_linalg ~op:"svd" (fun () -> caml_svd u s vt x);
  (u, s, vt)

external caml_eig :
  (Complex.t, Dtype.complex64_elt) t -> (Complex.t, Dtype.complex64_elt) t -> ('a, 'b) t -> bool -> unit = "nx_cuda_eig"

let eig_values x =
  let nd = Array.length x.shape in
  create_tensor x.context Dtype.Complex128 (Array.append (Array.sub x.shape 0 (nd - 2)) [| x.shape.(nd - 1) |])

let eigvals x =
  let w = eig_values x in
  lift_linalg ~op:"eigvals" (fun () -> caml_eig w w x false);
  w

let eig x =
  let w = eig_values x and v = create_tensor x.context Dtype.Complex128 x.shape in
  lift_linalg ~op:"eig" (fun () -> caml_eig w v x true);
  (w, v)

(* ---- step capture (beyond Backend_intf.S; see INTEGRATION.md section 2b). Between [capture_begin]
   and [capture_end] every op issued on the context is recorded into a CUDA graph instead of running;
   [graph_launch] replays the whole step with one launch. Tensors created inside the capture are the
   replay's outputs and keep their addresses; tensors that existed before it are read in place. ---- *)
type graph (* custom block around nxc_graph*, finalizer -> nxc_graph_destroy *)

external capture_begin : context -> unit = "nx_cuda_capture_begin"
external capture_end : context -> graph = "nx_cuda_capture_end"
external graph_launch : context -> graph -> unit = "nx_cuda_graph_launch"
external sync : context -> unit = "nx_cuda_sync"

let capture ctx f =
  capture_begin ctx;
  match f () with
  | v -> (v, capture_end ctx)
  | exception e -> (try ignore (capture_end ctx) with _ -> ()); raise e
