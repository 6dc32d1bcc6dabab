% tools/acados_golden.m — produce golden vectors from the REAL reference (MATLAB + acados v0.2.1 + CasADi), in the format
% tests/test_acados_golden.py loads.  This is the only route from "parity unpinned" to "parity pinned" (DESIGN.md 2.3):
% nothing of acados can run in this repository's build container.
%
% Run it in the reference workspace:
%     cd <uclv_qs_pushing_matlab>/acados_nmpc
%     source env.sh   (acados v0.2.1, README.md:21)            % ACADOS_INSTALL_DIR, MATLABPATH, LD_LIBRARY_PATH
%     python <this repo>/tools/acados_golden_inputs.py acados_golden_inputs.mat        % the seeded inputs, same as the tests'
%     matlab -batch "run('<this repo>/tools/acados_golden.m')"
%     cp acados_golden_*.mat <this repo>/tests/golden/acados/
% then `python -m pytest tests/test_acados_golden.py` compares the oracle with every recorded solve and, on a mismatch,
% flips the recalled-semantics switches of the oracle one at a time and names the flip that removes it.
%
% What is recorded per solve (all through the reference's own public API, nothing of its source is copied):
%   inputs : x0 (as handed to NMPC_controller.solve), index_time, y_ref (6 x T), warm start utraj / xtraj / ptraj before the call,
%            W, W_e, lh, uh, Hp, sample_time, object name, nlp_solver
%   outputs: u0 = controller.solve(...), and from controller.ocp_solver AFTER the call: get('x'), get('u'), get('pi'), get('status'),
%            get('sqp_iter'), get_cost, timing stats, get('stat') (the SQP iteration table: residuals, qp_iter, alpha) when available,
%            get('lam') / get('t') / get('sl') when the interface has them; controller.xtraj/utraj/ptraj after the call (shifted).
% nlp_solver: the reference hard-codes "sqp" (NMPC_controller.m:272).  For the SQP-RTI vectors change that literal to "sqp_rti",
% re-run with nlp = 'sqp_rti' below; both files are understood by the test.

nlp = 'sqp';                                  % must equal the literal at NMPC_controller.m:272 of the workspace that runs this
in = load('acados_golden_inputs.mat');        % written by tools/acados_golden_inputs.py

addpath(fullfile(pwd,'.')); addpath(fullfile(pwd,'..')); addpath(fullfile(pwd,'../cad_models')); addpath(fullfile(pwd,'./objects_database'));

cases = {};
for ic = 1:numel(in.cfg)
    cfg = in.cfg{ic};
    object_name = string(cfg.object);
    slider = object_selection(object_name);                                                            % main.m:26-27
    p = PusherSliderModel('real_plant', slider, 0, slider.cad_model_path, 3, slider.pcl_path, object_name);   % main.m:34
    p.symbolic_model_variable_shape();                                                                 % main.m:35
    controller = NMPC_controller('NMPC', p, cfg.dt, double(cfg.Hp));                                   % main.m:44
    controller.create_ocp_solver();                                                                    % main.m:45
    controller.set_delay_comp(0);
    controller.update_cost_function(cfg.W(1:4,1:4), cfg.W(5:6,5:6), cfg.We, 0, double(cfg.Hp)-1);      % main.m:82-86
    ocp = controller.ocp_solver;

    % ---- (a) independent single solves from given warm starts (config 3 shape)
    for k = 1:size(cfg.x0, 2)
        controller.initial_condition_update(cfg.x0(:,k));                                              % clears xtraj / utraj / ptraj
        controller.set_reference_trajectory(squeeze(cfg.y_ref(:,:,k)));
        if ~cfg.cold
            controller.xtraj = zeros(4, double(cfg.Hp)+1);
            controller.utraj = squeeze(cfg.u_init(:,:,k));
            controller.ptraj = zeros(4, double(cfg.Hp));
        end
        c = record_solve(controller, ocp, cfg, cfg.x0(:,k), 1, nlp, sprintf('%s_single_%d', cfg.name, k));
        c.y_ref = squeeze(cfg.y_ref(:,:,k));
        cases{end+1} = c; %#ok<SAGROW>
    end
    % ---- (b) closed loop = helper.closed_loop_matlab's loop body (helper.m:244-307), no noise / disturbance / delays (config 1)
    if cfg.closed_loop_steps > 0
        controller.initial_condition_update(cfg.cl_x0);
        controller.set_reference_trajectory(cfg.cl_y_ref);
        x = cfg.cl_x0;
        for i = 1:double(cfg.closed_loop_steps)
            c = record_solve(controller, ocp, cfg, x, i, nlp, sprintf('%s_loop_%d', cfg.name, i));
            c.y_ref = cfg.cl_y_ref;
            cases{end+1} = c; %#ok<SAGROW>
            x_dot = p.evalModelVariableShape(x, c.u0);                                                 % helper.m:294
            x = x + cfg.dt * x_dot;                                                                    % helper.m:307
        end
    end
    clear controller ocp
end
% the spline tables the reference built (pins a3-a8 of SURVEY.md 8a independently of the packaged tables)
save(sprintf('acados_golden_%s.mat', nlp), 'cases', 'nlp', '-v7');
fprintf('wrote acados_golden_%s.mat with %d recorded solves\n', nlp, numel(cases));

function c = record_solve(controller, ocp, cfg, x0, index_time, nlp, name)
    c = struct();
    c.name = name; c.nlp = nlp; c.object = cfg.object; c.Hp = double(cfg.Hp); c.dt = cfg.dt; c.W = cfg.W; c.We = cfg.We;
    lb = controller.h_constr_lb(:); ub = controller.h_constr_ub(:);
    c.lh = [-0.06; lb(2:end)]; c.uh = [0.011; ub(2:end)];                                              % the bounds the OCP is built with, NMPC_controller.m:251-252
    c.x0 = x0(:); c.index_time = index_time;
    c.first_call = isempty(controller.utraj) || isempty(controller.xtraj);
    c.utraj_in = controller.utraj; c.xtraj_in = controller.xtraj; c.ptraj_in = controller.ptraj;
    c.b = double(controller.plant.SP.b); c.knots = double(controller.plant.SP.S(:)); c.ctrl = double(controller.plant.SP.P);
    c.u0 = controller.solve(x0(:), index_time);                                                        % NMPC_controller.m:329
    c.x = ocp.get('x'); c.u = ocp.get('u'); c.pi = ocp.get('pi');
    c.status = ocp.get('status'); c.sqp_iter = ocp.get('sqp_iter'); c.cost = ocp.get_cost();
    c.time_tot = ocp.get('time_tot'); c.time_lin = ocp.get('time_lin'); c.time_qp_sol = ocp.get('time_qp_sol');
    c.xtraj_out = controller.xtraj; c.utraj_out = controller.utraj; c.ptraj_out = controller.ptraj;
    opt = {'stat', 'lam', 't', 'qp_iter', 'residuals'};
    for j = 1:numel(opt)
        try, c.(opt{j}) = ocp.get(opt{j}); catch, c.(opt{j}) = []; end %#ok<CTCH>
    end
    try, [c.v_bound_x0, c.t_angle_x0] = controller.update_tangential_velocity_bounds(x0(4)); catch, c.v_bound_x0 = []; c.t_angle_x0 = []; end %#ok<CTCH>
end
