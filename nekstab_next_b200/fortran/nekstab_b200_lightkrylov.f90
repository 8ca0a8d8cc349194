!> LightKrylov adapter (SURVEY.md section 8, row f-1): the device vector and the device operator as
!> extensions of LightKrylov's abstract types, with exactly the type-bound procedures the reference
!> itself implements for them:
!>   real_nek_vector  : zero, dot, scal, axpby            (core/nek_vectors.f90:20-31, 70-139)
!>   exponential_prop : matvec, rmatvec                   (core/linear_operators.f90:17-23, 39-44)
!> so that linear_stability_analysis / transient_growth_analysis (core/linear_stab.f90:12-119) compile
!> against device-resident vectors by changing two type names.
!>
!> STATUS: written against those in-tree signatures; NOT COMPILED in this repository's image (no Fortran
!> compiler, LightKrylov is not vendored).  Through this adapter LightKrylov's own Arnoldi still issues one
!> dot and one axpby per basis vector (2k sweeps and 2k all-reduces per step); the fused path is
!> eigs_d / svds_d / arnoldi_factorization_d of module nekstab_b200, which keep the whole loop on the device.
module nekstab_b200_lightkrylov
   use, intrinsic :: iso_c_binding
   use LightKrylov
   use nekstab_b200
   implicit none
   private

   !> A (basis, column) handle.  No default initialisation on purpose: LightKrylov passes vectors as
   !> intent(out) (core/linear_operators.f90:43), which would reset default-initialised components.
   type, extends(abstract_vector), public :: nek_dvector_lk
      type(c_ptr) :: basis
      integer(c_int) :: col
   contains
      private
      procedure, pass(self), public :: zero => lk_zero
      procedure, pass(self), public :: dot => lk_dot
      procedure, pass(self), public :: scal => lk_scal
      procedure, pass(self), public :: axpby => lk_axpby
   end type nek_dvector_lk

   !> matvec / rmatvec = nsb_op_apply of two operator handles (host callbacks wrapping the reference's
   !> direct / adjoint time-steppers, or device operators such as nsb_op_create_stepper).
   type, extends(abstract_linop), public :: device_linop
      type(c_ptr) :: op, op_adj
      real(c_double) :: t                     !< integration time, as exponential_prop%t (core/linear_stab.f90:72)
   contains
      private
      procedure, pass(self), public :: matvec => lk_matvec
      procedure, pass(self), public :: rmatvec => lk_rmatvec
   end type device_linop

   public :: device_basis_vectors

contains

   !> X(i) <- column i-1 of a device basis: replaces allocate(X(k_dim + 1)) (core/linear_stab.f90:60).
   subroutine device_basis_vectors(basis, X)
      type(c_ptr), intent(in) :: basis
      type(nek_dvector_lk), intent(inout) :: X(:)
      integer :: i
      do i = 1, size(X)
         X(i)%basis = basis
         X(i)%col = int(i - 1, c_int)
      end do
   end subroutine device_basis_vectors

   subroutine lk_zero(self)
      class(nek_dvector_lk), intent(inout) :: self
      call nsb_check(nsb_vec_zero(self%basis, self%col), 'zero')
   end subroutine lk_zero

   real(c_double) function lk_dot(self, vec) result(alpha)
      class(nek_dvector_lk), intent(in) :: self
      class(abstract_vector), intent(in) :: vec
      alpha = 0.0_c_double
      select type (vec)
      type is (nek_dvector_lk)
         ! BM1-weighted over vx, vy, [vz], t (pressure never) plus time*time, NaN -> error code -> nek_end
         call nsb_check(nsb_vec_dot(self%basis, self%col, vec%basis, vec%col, alpha), 'dot')
      end select
   end function lk_dot

   subroutine lk_scal(self, alpha)
      class(nek_dvector_lk), intent(inout) :: self
      real(c_double), intent(in) :: alpha
      call nsb_check(nsb_vec_scal(self%basis, self%col, alpha), 'scal')
   end subroutine lk_scal

   !> self <- alpha*self + beta*vec; %time untouched, like real_axpby (core/nek_vectors.f90:127-139)
   subroutine lk_axpby(self, alpha, vec, beta)
      class(nek_dvector_lk), intent(inout) :: self
      class(abstract_vector), intent(in) :: vec
      real(c_double), intent(in) :: alpha, beta
      select type (vec)
      type is (nek_dvector_lk)
         call nsb_check(nsb_vec_axpby(self%basis, self%col, alpha, vec%basis, vec%col, beta, NSB_AXPBY_SKIP_TIME), &
                        'axpby')
      end select
   end subroutine lk_axpby

   subroutine lk_matvec(self, vec_in, vec_out)
      class(device_linop), intent(in) :: self
      class(abstract_vector), intent(in) :: vec_in
      class(abstract_vector), intent(out) :: vec_out
      select type (vec_in)
      type is (nek_dvector_lk)
         select type (vec_out)
         type is (nek_dvector_lk)
            call nsb_check(nsb_op_apply(self%op, vec_in%basis, vec_in%col, vec_out%basis, vec_out%col), 'matvec')
         end select
      end select
   end subroutine lk_matvec

   subroutine lk_rmatvec(self, vec_in, vec_out)
      class(device_linop), intent(in) :: self
      class(abstract_vector), intent(in) :: vec_in
      class(abstract_vector), intent(out) :: vec_out
      select type (vec_in)
      type is (nek_dvector_lk)
         select type (vec_out)
         type is (nek_dvector_lk)
            call nsb_check(nsb_op_apply(self%op_adj, vec_in%basis, vec_in%col, vec_out%basis, vec_out%col), 'rmatvec')
         end select
      end select
   end subroutine lk_rmatvec

end module nekstab_b200_lightkrylov
