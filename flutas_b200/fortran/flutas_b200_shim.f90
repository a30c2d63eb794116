!
! flutas_b200_shim.f90 -- thin iso_c_binding shim between FluTAS's Fortran call sites and libflutas_b200.so.
!
! Drop this file into src/, list it in src/Makefile before main__<APP>.o and link with -lflutas_b200.
! It provides module procedures with the reference's names and argument lists, so the RK loop in
! src/apps/<APP>/main__<APP>.f90 (:707-726), initsolver.f90 and the BC-driven transform selection stay
! unchanged:
!     use mod_fft       -> fftini, fftend        (replaces src/fft.f90:24,159 ; initsolver.f90:117 calls fftini)
!     use mod_solver_gpu-> solver_gpu            (replaces src/solver_gpu.f90:31 ; same for solver_cpu)
!     use mod_fillps    -> fillps                (replaces src/fillps.f90:16)
!     use mod_correc    -> correc                (replaces src/correc.f90:16)
!     use mod_chkdiv    -> chkdiv                (replaces src/chkdiv.f90:18)
!     use mod_bound_b200-> boundp                (replaces src/bound.f90:146 for nh_p = 1, i.e. p and pold)
! This image has no Fortran compiler, so the file is syntax-simple F2003 and is not built here; the same
! C entry points are exercised through ctypes by tests/ (see INTEGRATION.md).
!
module mod_flutas_b200
  use, intrinsic :: iso_c_binding
  implicit none
  interface
    integer(c_int) function flutas_b200_init(device,rank,nranks) bind(C,name='flutas_b200_init')
      import; integer(c_int), value :: device,rank,nranks
    end function
    integer(c_int) function flutas_b200_fftini(n_x,n_y,bcxy,c_or_f,arrplan,normfft) bind(C,name='flutas_b200_fftini')
      import; integer(c_int), intent(in) :: n_x(3),n_y(3)
      character(kind=c_char), intent(in) :: bcxy(4),c_or_f(2)
      type(c_ptr), intent(out) :: arrplan(4); real(c_double), intent(out) :: normfft
    end function
    integer(c_int) function flutas_b200_fftend(arrplan) bind(C,name='flutas_b200_fftend')
      import; type(c_ptr), intent(inout) :: arrplan(4)
    end function
    integer(c_int) function flutas_b200_solver(n,arrplan,normfft,lambdaxy,a,b,c,bcz,c_or_f,p) bind(C,name='flutas_b200_solver')
      import; integer(c_int), intent(in) :: n(3); type(c_ptr), intent(in) :: arrplan(4)
      real(c_double), value :: normfft; type(c_ptr), value :: lambdaxy,a,b,c,p
      character(kind=c_char), intent(in) :: bcz(2),c_or_f(3)
    end function
    integer(c_int) function flutas_b200_solver_slab(n,arrplan,normfft,lambdaxy_g,a,b,c,bcz,c_or_f,p) &
                            bind(C,name='flutas_b200_solver_slab')
      import; integer(c_int), intent(in) :: n(3); type(c_ptr), intent(in) :: arrplan(4)
      real(c_double), value :: normfft; type(c_ptr), value :: lambdaxy_g,a,b,c,p
      character(kind=c_char), intent(in) :: bcz(2),c_or_f(3)
    end function
    integer(c_int) function flutas_b200_fillps(nx,ny,nz,nh_d,nh_u,dxi,dyi,dzi,dzfi,dti,rho0,u,v,w,p) &
                            bind(C,name='flutas_b200_fillps')
      import; integer(c_int), value :: nx,ny,nz,nh_d,nh_u; real(c_double), value :: dxi,dyi,dzi,dti,rho0
      type(c_ptr), value :: dzfi,u,v,w,p
    end function
    integer(c_int) function flutas_b200_correc(nx,ny,nz,nh_d,nh_u,dxi,dyi,dzi,dzci,dt,rho0,p,u,v,w,rho) &
                            bind(C,name='flutas_b200_correc')
      import; integer(c_int), value :: nx,ny,nz,nh_d,nh_u; real(c_double), value :: dxi,dyi,dzi,dt,rho0
      type(c_ptr), value :: dzci,p,u,v,w,rho
    end function
    integer(c_int) function flutas_b200_chkdiv(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,dzfi,u,v,w,divtot,divmax) &
                            bind(C,name='flutas_b200_chkdiv')
      import; integer(c_int), value :: nx,ny,nz,nh_d,nh_u; real(c_double), value :: dxi,dyi,dzi
      type(c_ptr), value :: dzfi,u,v,w; real(c_double), intent(out) :: divtot,divmax
    end function
    integer(c_int) function flutas_b200_pres_sp_src(nx,ny,nz,f_t12,dxi,dyi,dzi,nh_d,nh_u,dzci,rho0i,pold,u,v,w) &
                            bind(C,name='flutas_b200_pres_sp_src')
      import; integer(c_int), value :: nx,ny,nz,nh_d,nh_u; real(c_double), value :: f_t12,dxi,dyi,dzi,rho0i
      type(c_ptr), value :: dzci,pold,u,v,w
    end function
    integer(c_int) function flutas_b200_pres_tw_src(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,dzci,rho0i,f_t12,f_t12_o,p,pold,rho,u,v,w) &
                            bind(C,name='flutas_b200_pres_tw_src')
      import; integer(c_int), value :: nx,ny,nz,nh_d,nh_u; real(c_double), value :: dxi,dyi,dzi,rho0i,f_t12,f_t12_o
      type(c_ptr), value :: dzci,p,pold,rho,u,v,w
    end function
    integer(c_int) function flutas_b200_load(io,filename,ng,n,start,nh,fld) bind(C,name='flutas_b200_load')
      import; character(kind=c_char), value :: io; character(kind=c_char), intent(in) :: filename(*)
      integer(c_int), intent(in) :: ng(3),n(3),start(3); integer(c_int), value :: nh; type(c_ptr), value :: fld
    end function
    integer(c_int) function flutas_b200_pold_update(nx,ny,nz,mode,p,pold) bind(C,name='flutas_b200_pold_update')
      import; integer(c_int), value :: nx,ny,nz,mode; type(c_ptr), value :: p,pold
    end function
    integer(c_int) function flutas_b200_boundp(cbc,n,bc,nh_d,nh_p,dl,dzc,dzf,p) bind(C,name='flutas_b200_boundp')
      import; character(kind=c_char), intent(in) :: cbc(6); integer(c_int), intent(in) :: n(3)
      real(c_double), intent(in) :: bc(6),dl(3); integer(c_int), value :: nh_d,nh_p
      type(c_ptr), value :: dzc,dzf,p
    end function
    function flutas_b200_last_error() bind(C,name='flutas_b200_last_error') result(msg)
      import; type(c_ptr) :: msg
    end function
  end interface
contains
  subroutine b200_check(istat,where)          ! the reference prints and stops on errors (src/fft.f90:879-883)
    integer(c_int), intent(in) :: istat
    character(len=*), intent(in) :: where
    if(istat.ne.0) then
      print*, 'flutas_b200 error in ', where, ' (see flutas_b200_last_error)'
      error stop 1
    endif
  end subroutine b200_check
end module mod_flutas_b200
!
module mod_fft                                ! same public names as src/fft.f90:17
  use, intrinsic :: iso_c_binding
  use mod_flutas_b200
  use mod_types
  implicit none
  private
  public :: fftini,fftend
contains
  subroutine fftini(n_x,n_y,bcxy,c_or_f,arrplan,normfft)
    integer         , intent(in ), dimension(3)     :: n_x,n_y
    character(len=1), intent(in ), dimension(0:1,2) :: bcxy
    character(len=1), intent(in ), dimension(2)     :: c_or_f
    type(C_PTR)     , intent(out), dimension(2,2)   :: arrplan
    real(rp)        , intent(out)                   :: normfft
    type(C_PTR) :: plans(4)
    character(kind=c_char) :: cb(4),cf(2)
    cb = (/bcxy(0,1),bcxy(1,1),bcxy(0,2),bcxy(1,2)/)
    cf = (/c_or_f(1),c_or_f(2)/)
    call b200_check(flutas_b200_fftini(int(n_x,c_int),int(n_y,c_int),cb,cf,plans,normfft),'fftini')
    arrplan = reshape(plans,(/2,2/))           ! (1,1) fwd-x (2,1) bwd-x (1,2) fwd-y (2,2) bwd-y, as fft.f90:151-154
  end subroutine fftini
  subroutine fftend(arrplan)
    type(C_PTR), intent(in), dimension(2,2) :: arrplan
    type(C_PTR) :: plans(4)
    plans = reshape(arrplan,(/4/))
    call b200_check(flutas_b200_fftend(plans),'fftend')
  end subroutine fftend
end module mod_fft
!
module mod_solver_gpu                         ! same signature as src/solver_gpu.f90:31-47
  use, intrinsic :: iso_c_binding
  use mod_flutas_b200
  use mod_common_mpi, only: n_z
  use mod_types
  implicit none
  private
  public :: solver_gpu
contains
  subroutine solver_gpu(n,dims,arrplan,normfft,lambdaxy,a,b,c,bc,c_or_f,p)
    integer         , intent(in   ), dimension(3)             :: n,dims
    type(C_PTR)     , intent(in   ), dimension(2,2)           :: arrplan
    real(rp)        , intent(in   )                           :: normfft
    real(rp)        , intent(in   ), dimension(:,:), target   :: lambdaxy
    real(rp)        , intent(in   ), dimension(:)  , target   :: a,b,c
    character(len=1), intent(in   ), dimension(0:1,3)         :: bc
    character(len=1), intent(in   ), dimension(3)             :: c_or_f
    real(rp)        , intent(inout), dimension(0:,0:,0:), target :: p
    type(C_PTR) :: plans(4)
    character(kind=c_char) :: bz(2),cf(3)
    plans = reshape(arrplan,(/4/))
    bz = (/bc(0,3),bc(1,3)/); cf = c_or_f
    if(dims(1)*dims(2).eq.1) then
      call b200_check(flutas_b200_solver(int(n,c_int),plans,normfft,c_loc(lambdaxy),c_loc(a),c_loc(b),c_loc(c), &
                                         bz,cf,c_loc(p)),'solver')
    else
      ! slab decomposition dims_in = (1,nranks): `lambdaxy` must hold the all-gathered (ng1,ng2) eigenvalues
      ! (one MPI_ALLGATHER of the initsolver windows along y, done once after initsolver; INTEGRATION.md)
      call b200_check(flutas_b200_solver_slab(int(n,c_int),plans,normfft,c_loc(lambdaxy),c_loc(a),c_loc(b),c_loc(c), &
                                              bz,cf,c_loc(p)),'solver_slab')
    endif
  end subroutine solver_gpu
end module mod_solver_gpu
!
module mod_fillps                             ! same signature as src/fillps.f90:16-26
  use, intrinsic :: iso_c_binding
  use mod_flutas_b200
  use mod_types
  implicit none
  private
  public :: fillps
contains
  subroutine fillps(nx,ny,nz,nh_d,nh_u,dxi,dyi,dzi,dzfi,dti,rho0,u,v,w,p)
    integer , intent(in )                                     :: nx,ny,nz,nh_d,nh_u
    real(rp), intent(in )                                     :: dxi,dyi,dzi,dti,rho0
    real(rp), intent(in ), dimension(1-nh_d:), target         :: dzfi
    real(rp), intent(in ), dimension(1-nh_u:,1-nh_u:,1-nh_u:), target :: u,v,w
    real(rp), intent(out), dimension(0:,0:,0:), target        :: p
    call b200_check(flutas_b200_fillps(nx,ny,nz,nh_d,nh_u,dxi,dyi,dzi,c_loc(dzfi),dti,rho0, &
                                       c_loc(u),c_loc(v),c_loc(w),c_loc(p)),'fillps')
  end subroutine fillps
end module mod_fillps
!
module mod_correc                             ! same signature as src/correc.f90:16-29
  use, intrinsic :: iso_c_binding
  use mod_flutas_b200
  use mod_types
  implicit none
  private
  public :: correc
contains
  subroutine correc(nx,ny,nz,nh_d,nh_u,dxi,dyi,dzi,dzci,dt,rho0,p,u,v,w,rho)
    integer , intent(in   )                                     :: nx,ny,nz,nh_d,nh_u
    real(rp), intent(in   )                                     :: dxi,dyi,dzi,dt,rho0
    real(rp), intent(in   ), dimension(1-nh_d:), target         :: dzci
    real(rp), intent(in   ), dimension(0:,0:,0:), target        :: p
    real(rp), intent(inout), dimension(1-nh_u:,1-nh_u:,1-nh_u:), target :: u,v,w
    real(rp), intent(in   ), dimension(0:,0:,0:)                :: rho      ! (0,0,0)-sized dummy: never touched
    call b200_check(flutas_b200_correc(nx,ny,nz,nh_d,nh_u,dxi,dyi,dzi,c_loc(dzci),dt,rho0, &
                                       c_loc(p),c_loc(u),c_loc(v),c_loc(w),c_null_ptr),'correc')
  end subroutine correc
end module mod_correc
!
module mod_chkdiv                             ! same signature as src/chkdiv.f90:18-29
  use, intrinsic :: iso_c_binding
  use mpi
  use mod_flutas_b200
  use mod_common_mpi, only: myid,ierr
  use mod_types
  implicit none
  private
  public :: chkdiv
contains
  subroutine chkdiv(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,dzfi,u,v,w,divtot,divmax)
    integer , intent(in )                                     :: nx,ny,nz,nh_d,nh_u
    real(rp), intent(in )                                     :: dxi,dyi,dzi
    real(rp), intent(in ), dimension(1-nh_d:), target         :: dzfi
    real(rp), intent(in ), dimension(1-nh_u:,1-nh_u:,1-nh_u:), target :: u,v,w
    real(rp), intent(out)                                     :: divtot,divmax
    call b200_check(flutas_b200_chkdiv(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,c_loc(dzfi),c_loc(u),c_loc(v),c_loc(w), &
                                       divtot,divmax),'chkdiv')
    call mpi_allreduce(MPI_IN_PLACE,divtot,1,MPI_REAL_RP,MPI_SUM,MPI_COMM_WORLD,ierr)   ! as chkdiv.f90:64-65
    call mpi_allreduce(MPI_IN_PLACE,divmax,1,MPI_REAL_RP,MPI_MAX,MPI_COMM_WORLD,ierr)
    if(myid.eq.0) print*, 'Total divergence = ', divtot, '| Maximum divergence = ', divmax
  end subroutine chkdiv
end module mod_chkdiv

!
! boundp for the pressure halo (nh_p = 1): same argument list as src/bound.f90:146; `halo` (MPI datatypes) is unused.
! cbc(0:1,3) and bc(0:1,3) are passed in Fortran storage order = (x0,x1,y0,y1,z0,z1), which is what the C side expects.
!
module mod_bound_b200
  use, intrinsic :: iso_c_binding
  use mod_flutas_b200
  implicit none
  private
  public :: boundp
  contains
  subroutine boundp(cbc,n,bc,nh_d,nh_p,halo,dl,dzc,dzf,p)
    character(len=1), intent(in   ), dimension(0:1,3)                   :: cbc
    integer         , intent(in   ), dimension(3)                       :: n
    real(c_double)  , intent(in   ), dimension(0:1,3)                   :: bc
    integer         , intent(in   )                                     :: nh_d,nh_p
    integer         , intent(in   ), dimension(3)                       :: halo
    real(c_double)  , intent(in   ), dimension(3)                       :: dl
    real(c_double)  , intent(in   ), dimension(1-nh_d:), target         :: dzc,dzf
    real(c_double)  , intent(inout), dimension(1-nh_p:,1-nh_p:,1-nh_p:), target :: p
    character(kind=c_char) :: cc(6)
    real(c_double) :: bv(6)
    integer :: d,s
    do d=1,3
      do s=0,1
        cc(2*(d-1)+s+1) = cbc(s,d)
        bv(2*(d-1)+s+1) = bc(s,d)
      enddo
    enddo
    call b200_check(flutas_b200_boundp(cc,int(n,c_int),bv,int(nh_d,c_int),int(nh_p,c_int),dl, &
                                       c_loc(dzc),c_loc(dzf),c_loc(p)),'boundp')
  end subroutine boundp
end module mod_bound_b200

!
! pressure-gradient source terms of the predictor, same signatures as src/source.f90:247,311 (constant-coefficient
! Poisson build), and the two pressure bookkeeping loops of main__single_phase.f90:693-699 / :734-740 as
!   call pold_update(n,0,p,pold)   ! pold = p        call pold_update(n,1,p,pold)   ! p = pold + p
!
module mod_source_b200
  use, intrinsic :: iso_c_binding
  use mod_flutas_b200
  implicit none
  private
  public :: pres_sp_src,pres_tw_src,pold_update
  contains
  subroutine pres_sp_src(nx,ny,nz,f_t12,dxi,dyi,dzi,nh_d,nh_u,dzci,rho0i,pold,u,v,w)
    integer       , intent(in   )                                     :: nx,ny,nz,nh_d,nh_u
    real(c_double), intent(in   )                                     :: f_t12,dxi,dyi,dzi,rho0i
    real(c_double), intent(in   ), dimension(1-nh_d:), target         :: dzci
    real(c_double), intent(in   ), dimension(0:,0:,0:), target        :: pold
    real(c_double), intent(inout), dimension(1-nh_u:,1-nh_u:,1-nh_u:), target :: u,v,w
    call b200_check(flutas_b200_pres_sp_src(nx,ny,nz,f_t12,dxi,dyi,dzi,nh_d,nh_u,c_loc(dzci),rho0i, &
                                            c_loc(pold),c_loc(u),c_loc(v),c_loc(w)),'pres_sp_src')
  end subroutine pres_sp_src
  subroutine pres_tw_src(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,dzci,rho0i,f_t12,f_t12_o,p,pold,rho,u,v,w)
    integer       , intent(in   )                                     :: nx,ny,nz,nh_d,nh_u
    real(c_double), intent(in   )                                     :: dxi,dyi,dzi,rho0i,f_t12,f_t12_o
    real(c_double), intent(in   ), dimension(1-nh_d:), target         :: dzci
    real(c_double), intent(in   ), dimension(0:,0:,0:), target        :: p,pold,rho
    real(c_double), intent(inout), dimension(1-nh_u:,1-nh_u:,1-nh_u:), target :: u,v,w
    call b200_check(flutas_b200_pres_tw_src(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,c_loc(dzci),rho0i,f_t12,f_t12_o, &
                                            c_loc(p),c_loc(pold),c_loc(rho),c_loc(u),c_loc(v),c_loc(w)),'pres_tw_src')
  end subroutine pres_tw_src
  subroutine pold_update(n,mode,p,pold)
    integer       , intent(in   ), dimension(3)                 :: n
    integer       , intent(in   )                               :: mode
    real(c_double), intent(inout), dimension(0:,0:,0:), target  :: p,pold
    call b200_check(flutas_b200_pold_update(n(1),n(2),n(3),mode,c_loc(p),c_loc(pold)),'pold_update')
  end subroutine pold_update
end module mod_source_b200

!
! restart files: same signature as src/load.f90:21 (decomposition from 2DECOMP's xstart, global size from mod_param)
!
module mod_load_b200
  use, intrinsic :: iso_c_binding
  use mod_flutas_b200
  use mod_param , only: ng
  use decomp_2d , only: xstart
  implicit none
  private
  public :: load
  contains
  subroutine load(io,filename,n,fld)
    character(len=1), intent(in   )                                    :: io
    character(len=*), intent(in   )                                    :: filename
    integer         , intent(in   ), dimension(3)                      :: n
    real(c_double)  , intent(inout), dimension(n(1),n(2),n(3)), target :: fld
    call b200_check(flutas_b200_load(io,trim(filename)//c_null_char,int(ng,c_int),int(n,c_int),int(xstart-1,c_int), &
                                     0_c_int,c_loc(fld)),'load')
  end subroutine load
end module mod_load_b200
