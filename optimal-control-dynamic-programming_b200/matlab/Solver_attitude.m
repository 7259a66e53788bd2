classdef Solver_attitude < handle
    %SOLVER_ATTITUDE  Three (w, theta) axes with three torque levels, B200 back end.
    %   Drop-in for simplified_run of the reference class (attitude-control/Solver_attitude.m
    %   :196-259, :625-667).  The full 6-D run() of the reference never executed (SURVEY 2.3) and
    %   is not provided; the forward simulations are outside this build's scope.

    properties
        w_min
        w_max
        n_mesh_w
        yaw_min
        yaw_max
        pitch_min
        pitch_max
        roll_min
        roll_max
        n_mesh_q
        n_mesh_t
        InertiaM
        J1
        J2
        J3
        Q1
        Q2
        Q3
        Q4
        Q5
        Q6
        R1
        R2
        R3
        Qt1
        Qt2
        Qt3
        T_final
        h
        N_stage
        defaultX0
        U_vector
        U1_Opt
        U2_Opt
        U3_Opt
        F_Values
        U_idx
        device = -1
        last_desc_      % descriptor of the last simplified_run (for the forward simulation)
        policy6_        % grids and policy of the last run() (for get_optimal_path)
        defaultX0
        n_gpus = 1      % > 1: the grid is cut into slabs over this many GPUs, driven from this one process
    end

    methods
        function this = Solver_attitude()
            this.w_min = -deg2rad(50);  this.w_max = -deg2rad(-50);  this.n_mesh_w = 1000;
            this.yaw_min = -30;  this.yaw_max = 30;
            this.pitch_min = -20;  this.pitch_max = 20;
            this.roll_min = -35;  this.roll_max = 35;
            this.n_mesh_q = 10;  this.n_mesh_t = 300;
            i1 = 0.02836 + 0.00016; i2 = 0.026817 + 0.00150; i3 = 0.023 + 0.00150;
            i4 = -0.0000837; i5 = 0.000014; i6 = -0.00029;
            this.InertiaM = [i1 i4 i5; i4 i2 i6; i5 i6 i3];
            this.Q1 = 6; this.Q2 = 6; this.Q3 = 6; this.Q4 = 6; this.Q5 = 6; this.Q6 = 6;
            this.R1 = 4; this.R2 = 4; this.R3 = 4;
            this.Qt1 = this.Q4; this.Qt2 = this.Q5; this.Qt3 = this.Q6;
            this.T_final = 30;  this.h = 0.005;
            this.N_stage = ceil(this.T_final/this.h);
            this.J1 = this.InertiaM(1); this.J2 = this.InertiaM(5); this.J3 = this.InertiaM(9);
            this.U_vector = [-0.11 0 0.11];
            q0 = [0.0501511024391496;0.0833950587800888;-0.0818761044636256;0.991880252153991];   % :306-309
            this.defaultX0 = [0;0;0;q0];
        end

        function simplified_run(obj, n_stages)
            obj.N_stage = ceil(obj.T_final/obj.h);
            if nargin < 2, n_stages = obj.N_stage - 1; end
            s_w = linspace(obj.w_min, obj.w_max, obj.n_mesh_w).';
            lim = [obj.yaw_min obj.yaw_max; obj.pitch_min obj.pitch_max; obj.roll_min obj.roll_max];
            s_t = zeros(obj.n_mesh_t, 3);
            for a = 1:3, s_t(:,a) = linspace(deg2rad(lim(a,1)), deg2rad(lim(a,2)), obj.n_mesh_t).'; end
            U = obj.U_vector(:);  hh = obj.h;  Jax = [obj.J1 obj.J2 obj.J3];
            k1 = s_w; k2 = s_w + k1*hh/2; k3 = s_w + k2*hh/2; k4 = s_w + k3*hh;
            inct = hh*(k1 + 2*k2 + 2*k3 + k4)/6;             % theta' = T + inct(w)
            incw = zeros(numel(U), 3);
            for a = 1:3, kk = U/Jax(a); incw(:,a) = hh*(kk + 2*kk + 2*kk + kk)/6; end
            Qw = [obj.Q1 obj.Q2 obj.Q3]; Qt = [obj.Qt1 obj.Qt2 obj.Qt3]; R = [obj.R1 obj.R2 obj.R3];
            rep = @(v) repmat(v, 1, 3);
            d.n = [obj.n_mesh_w, obj.n_mesh_t];  d.C = numel(U);  d.P = 3;  d.N = obj.N_stage;
            d.grid = {rep(s_w), s_t};
            d.src_a = [1 2];  d.src_b = [0 1];  d.q_order = [1 2];
            d.Ta = {rep(s_w), s_t};
            d.Tb = {[], rep(inct)};
            d.Tc = {incw, []};
            qt = zeros(size(s_t)); for a = 1:3, qt(:,a) = Qt(a)*s_t(:,a).^2; end
            d.q  = {(s_w.^2)*Qw, qt};
            d.r  = (U.^2)*R;
            d.store_J_all = 0; d.store_idx_all = 0; d.device = obj.device;
            obj.last_desc_ = d;
            sz = [obj.n_mesh_w, obj.n_mesh_t, 3];
            tic
            if obj.n_gpus > 1
                [Jv, iv] = bellman_sweep_multi(d, n_stages, obj.n_gpus, struct());
                obj.F_Values = reshape(Jv, sz);  obj.U_idx = double(reshape(iv, sz));
            else
                hnd = bellman_mex('create', d);
                bellman_mex('run', hnd, n_stages, struct());
                obj.F_Values = reshape(bellman_mex('get_J', hnd), sz);
                obj.U_idx = double(reshape(bellman_mex('get_idx', hnd), sz));
                bellman_mex('destroy', hnd);
            end
            fprintf('%d stages - %f seconds\n', n_stages, toc)
            obj.U1_Opt = griddedInterpolant({s_w.', s_t(:,1).'}, obj.U_vector(obj.U_idx(:,:,1)), 'nearest');
            obj.U2_Opt = griddedInterpolant({s_w.', s_t(:,2).'}, obj.U_vector(obj.U_idx(:,:,2)), 'nearest');
            obj.U3_Opt = griddedInterpolant({s_w.', s_t(:,3).'}, obj.U_vector(obj.U_idx(:,:,3)), 'nearest');
            fprintf('...Done!\n')
        end

        function run(obj, n_stages)
            % Solver_attitude.run of the reference (:521-601): the coupled sweep over (w1 w2 w3 yaw pitch
            % roll) with 27 control combinations.  The next-state arrays are built here exactly as
            % reshape_states / spacecraft_dynamics_taylor_estimate do (implicit expansion), but each only
            % over the dimensions it depends on; the repmat to nine dimensions, J_current_state_fix + F(...)
            % and the three nested min calls are one fused GPU stage (bellman_mex('dense6_run', ...)).
            % The reference's default n_mesh_w = 1000 cannot be run anywhere (2.7e13-element arrays):
            % set n_mesh_w / n_mesh_q to a mesh that fits (9 doubles per state on the host).
            obj.N_stage = ceil(obj.T_final/obj.h);
            if nargin < 2, n_stages = obj.N_stage - 1; end
            nw = obj.n_mesh_w;  nq = obj.n_mesh_q;  U = obj.U_vector(:);  nu = numel(U);  hh = obj.h;
            sr = linspace(obj.w_min, obj.w_max, nw);
            s_yaw = linspace(deg2rad(obj.yaw_min), deg2rad(obj.yaw_max), nq);
            s_pitch = linspace(deg2rad(obj.pitch_min), deg2rad(obj.pitch_max), nq);
            s_roll = linspace(deg2rad(obj.roll_min), deg2rad(obj.roll_max), nq);
            X1V = reshape(sr, [nw 1]);  X2V = reshape(sr, [1 nw]);  X3V = reshape(sr, [1 1 nw]);
            c4 = reshape(cos(s_yaw/2), [1 1 1 nq]);      s4 = reshape(sin(s_yaw/2), [1 1 1 nq]);
            c5 = reshape(cos(s_pitch/2), [1 1 1 1 nq]);  s5 = reshape(sin(s_pitch/2), [1 1 1 1 nq]);
            c6 = reshape(cos(s_roll/2), [1 1 1 1 1 nq]); s6 = reshape(sin(s_roll/2), [1 1 1 1 1 nq]);
            qa = s4.*c5.*c6 - c4.*s5.*s6;  qb = c4.*s5.*c6 + s4.*c5.*s6;  qc = c4.*c5.*s6 - s4.*s5.*c6;
            full = [nw nw nw nq nq nq];
            gs = obj.Q1*X1V.^2 + obj.Q2*X2V.^2 + obj.Q3*X3V.^2 + obj.Q4*qa.^2 + obj.Q5*qb.^2 + obj.Q6*qc.^2;
            x7 = (1 - (qa.^2 + qb.^2 + qc.^2)).^0.5;
            U1V = reshape(U, [1 1 1 nu]);
            w1n = X1V + hh*((obj.J2-obj.J3)/obj.J1*X2V.*X3V + U1V/obj.J1);
            w2n = X2V + hh*((obj.J3-obj.J1)/obj.J2*X3V.*X1V + U1V/obj.J2);
            w3n = X3V + hh*((obj.J1-obj.J2)/obj.J3*X1V.*X2V + U1V/obj.J3);
            X4n = qa + hh*(0.5*(X3V.*qb - X2V.*qc + X1V.*x7));
            X5n = qb + hh*(0.5*(-X3V.*qa + X1V.*qc + X2V.*x7));
            X6n = qc + hh*(0.5*(X2V.*qa - X1V.*qb + X3V.*x7));
            x7 = x7 + hh*(0.5*(-X1V.*qa - X2V.*qb - X3V.*qc));
            Qs = sqrt(X4n.^2 + X5n.^2 + X6n.^2 + x7.^2);
            X4n = X4n./Qs;  X5n = X5n./Qs;  X6n = X6n./Qs;  x7 = x7./Qs;
            yaw_n = atan2(2.*(X6n.*X5n + x7.*X4n), x7.^2 + X6n.^2 - X5n.^2 - X4n.^2);
            pitch_n = asin(-2.*(X6n.*X4n - x7.*X5n));
            roll_n = atan2(2.*(X5n.*X4n + x7.*X6n), x7.^2 - X6n.^2 - X5n.^2 + X4n.^2);
            ex = @(a) reshape(a + zeros(full), [], 1);
            d6 = struct('n', full, 'nu', nu, 'device', obj.device);
            d6.grid = {sr(:), sr(:), sr(:), s_yaw(:), s_pitch(:), s_roll(:)};
            d6.w_next = {reshape(w1n + zeros([nw nw nw nu]), [], nu), reshape(w2n + zeros([nw nw nw nu]), [], nu), ...
                         reshape(w3n + zeros([nw nw nw nu]), [], nu)};
            d6.a_next = {ex(yaw_n), ex(pitch_n), ex(roll_n)};
            d6.gs = ex(gs);
            d6.r = {obj.R1*U.^2, obj.R2*U.^2, obj.R3*U.^2};
            tic
            [J, id, ms] = bellman_mex('dense6_run', d6, n_stages, []);
            fprintf('%d stages - %f seconds (device %.3f ms)\n', n_stages, toc, ms)
            obj.policy6_ = struct('d6', rmfield(d6, {'w_next', 'a_next', 'gs', 'r'}), 'id', id);   % for get_optimal_path
            id = double(id) - 1;
            obj.F_Values = reshape(J, full);
            obj.U1_Opt = single(reshape(obj.U_vector(floor(id/(nu*nu)) + 1), full));     % :561-563
            obj.U2_Opt = single(reshape(obj.U_vector(mod(floor(id/nu), nu) + 1), full));
            obj.U3_Opt = single(reshape(obj.U_vector(mod(id, nu) + 1), full));
            fprintf('...Done!\n')
        end

        function [X, U] = get_optimal_path(obj, X0)
            % Solver_attitude.get_optimal_path of the reference (:1487-1530) with method 'nearest': the 6-D
            % policy of run(), quat2angle per step, first-order ('taylor') plant step; every column of X0
            % (7 x batch, default obj.defaultX0) is one GPU thread.  X: 7 x N x batch, U: 3 x (N-1) x batch.
            if nargin < 2, X0 = obj.defaultX0; end
            if isempty(obj.policy6_), error('run(obj) must complete before get_optimal_path'); end
            N = obj.N_stage;
            [Xf, Uf] = bellman_mex('rollout_attitude6', obj.policy6_.d6, obj.policy6_.id, obj.U_vector(:), ...
                [obj.J1 obj.J2 obj.J3], obj.h, N - 1, X0);
            batch = size(X0, 2);
            X = reshape(Xf, 7, N, batch);  U = reshape(Uf, 3, N - 1, batch);
            v_plot = 0:obj.h:obj.T_final-obj.h;
            figure; hold on; grid on; plot(v_plot(1:N-1), U(:,:,1).', '--'); legend('u1','u2','u3'); xlabel('time (s)')
            figure; hold on; grid on; plot(v_plot(1:N), X(1:3,:,1).'); legend('w1','w2','w3')
        end

        function [X_ode45, U_ode45] = get_optimal_path_simplified_testode45(obj, X0)
            % Solver_attitude.m:1669-1705 of the reference for every column of X0 (7 x batch: w1 w2 w3 q1 q2
            % q3 q4; default obj.defaultX0): U(k) = U{k}_Opt(X(k), 2*asin(X(3+k))), then ode45 over one
            % stage on the full rigid-body plant (:1803-1849), one GPU thread per initial state.
            if nargin < 2, X0 = obj.defaultX0; end
            d = obj.last_desc_;
            hnd = bellman_mex('create', d);
            bellman_mex('set_stage', hnd, 1, reshape(obj.F_Values, [], 3), reshape(obj.U_idx, [], 3));
            N = obj.N_stage;
            o = struct('n_steps', N - 1, 'stride_out', 1, 'h', obj.h, 'rtol', 1e-3, 'atol', 1e-6, 'InertiaM', obj.InertiaM);
            [X, id] = bellman_mex('rollout_attitude', hnd, 1, o, obj.U_vector(:), X0);
            bellman_mex('destroy', hnd);
            batch = size(X0, 2);
            X_ode45 = permute(reshape(X, 7, N, batch), [2 1 3]);                 % N x 7 (x batch), as :1683
            U_ode45 = permute(reshape(obj.U_vector(id), 3, N - 1, batch), [2 1 3]);
            T_ode45 = (0:N-1)*obj.h;
            figure; hold on; grid on; plot(T_ode45, X_ode45(:,1:3,1)*180/pi); legend('w1','w2','w3')
            figure; hold on; grid on; plot(T_ode45(1:end-1), U_ode45(:,:,1), '--'); legend('u1','u2','u3')
        end
    end
end
